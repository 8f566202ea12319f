// Flash-style multi-head self-attention over the C'*N+1 channel-patch tokens, on
// tcgen05 tensor cores with TMEM accumulators (head_dim = 64 for every DiChaViT size).
// Replaces reference models/vit.py:126-141:  softmax(q k^T * hd^-0.5) v, which there
// materialises the [B,H,L,L] probabilities; here only O and the per-row log-sum-exp
// are written.
//
// Input  qkv  bf16 [B, L, 3*D]   (row = token, columns [q | k | v], head h at h*64)
// Output o    bf16 [B, L, D]     (head-major columns == reference's transpose(1,2).reshape)
//        lse2 fp32 [B, H, Lp]    (log2-domain: max + log2(sum) of scale*log2e-scaled scores;
//                                 Lp = L rounded up to 128 so tiles are 16-byte aligned)
//
// Forward kernel, one CTA per (128-query tile, head, image), 2 CTAs resident per SM:
//   warps 0-3  softmax warpgroup: thread t owns query row t (TMEM lane t)
//   warp  4    TMA producer (Q once, K/V tiles through a 2-stage ring)
//   warp  5    MMA issuer + TMEM owner
//   TMEM columns: S[0,128) fp32 | P[128,192) bf16x2 (A operand of the PV MMA) | PV[192,256)
// Per KV tile: S = Q K^T (SS MMA) -> online softmax in registers -> P to TMEM ->
// PV = P V (TS MMA, V consumed MN-major straight from its TMA tile) -> O = O*alpha + PV.
#include "common.cuh"
#include "host.h"

namespace dcv {

constexpr int kHd = 64;
constexpr int kTq = 128;
constexpr int kTk = 128;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnFwdParams {
  int B, L, H, D;
  int Lp;     // row stride of lse2 (L rounded up to 128)
  float sl2;  // softmax scale * log2(e)
  __nv_bfloat16* o;
  float* lse2;
};

constexpr int kFwdStages = 2;
constexpr int kFwdSmem = kTq * kHd * 2 + kFwdStages * 2 * kTk * kHd * 2 + 1024 + 128;

__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTq * kHd * 2;                       // kFwdStages x 16 KB
  uint8_t* sV = sK + kFwdStages * kTk * kHd * 2;            // kFwdStages x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kFwdStages * kTk * kHd * 2);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = bars + 1 + kFwdStages;
  uint64_t* s_full = bars + 1 + 2 * kFwdStages;
  uint64_t* p_full = s_full + 1;
  uint64_t* pv_full = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTq;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.L + kTk - 1) / kTk;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < kFwdStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tP = tmem_base + 128, tPV = tmem_base + 192;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTq * kHd * 2);
      tma_load_3d(sQ, &map_qkv, q_full, h * kHd, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kFwdStages;
        mbar_wait(&kv_empty[st], ((j / kFwdStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * kTk * kHd * 2);
        tma_load_3d(sK + st * (kTk * kHd * 2), &map_qkv, &kv_full[st], p.D + h * kHd, j * kTk, b);
        tma_load_3d(sV + st * (kTk * kHd * 2), &map_qkv, &kv_full[st], 2 * p.D + h * kHd, j * kTk, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kTq, kTk, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kTq, kHd, 0, 1);  // B = V, MN-major
      const uint64_t dq = make_desc_kmajor(smem_u32(sQ));
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      {
        const uint64_t dk = make_desc_kmajor(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k) umma_ss(tS, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit(s_full);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kFwdStages;
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint64_t dv = make_desc_mnmajor(smem_u32(sV + st * (kTk * kHd * 2)), 64 * 128);
#pragma unroll
        for (int k = 0; k < kTk / 16; ++k) umma_ts(tPV, tP + 8 * k, dv + 128 * k, idesc_pv, k ? 1u : 0u);
        umma_commit(pv_full);
        umma_commit(&kv_empty[st]);
        if (j + 1 < n_kv) {
          const int st1 = (j + 1) % kFwdStages;
          mbar_wait(&kv_full[st1], ((j + 1) / kFwdStages) & 1);
          tc_fence_after();
          const uint64_t dk = make_desc_kmajor(smem_u32(sK + st1 * (kTk * kHd * 2)));
#pragma unroll
          for (int k = 0; k < kHd / 16; ++k) umma_ss(tS, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
          umma_commit(s_full);
        }
      }
    }
  } else {
    // ------------------------------ softmax warpgroup ------------------------------
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int row = q0 + warp * 32 + lane;
    float m = -INFINITY, l = 0.f;
    float o[kHd];
#pragma unroll
    for (int i = 0; i < kHd; ++i) o[i] = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int kv0 = j * kTk;
      const bool tail = kv0 + kTk > p.L;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kTk / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tS + lane_base + c * 32, r);
        tmem_ld_wait();
        if (tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kv0 + c * 32 + i < p.L) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float m_new = fmaxf(m, mx * p.sl2);
      const float alpha = fast_exp2(m - m_new);
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < kTk / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tS + lane_base + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.sl2, -m_new));
          float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.sl2, -m_new));
          if (tail) {
            if (kv0 + c * 32 + 2 * i >= p.L) p0 = 0.f;
            if (kv0 + c * 32 + 2 * i + 1 >= p.L) p1 = 0.f;
          }
          sum += p0 + p1;
          pk[i] = pack_bf16(p0, p1);
        }
        tmem_st16(tP + lane_base + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
      l = l * alpha + sum;
      m = m_new;

      mbar_wait(pv_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < kHd / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tPV + lane_base + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(r[i]));
      }
      tc_fence_before();
    }

    if (row < p.L) {
      const float inv = 1.0f / l;
      __nv_bfloat16* dst = p.o + (static_cast<size_t>(b) * p.L + row) * p.D + h * kHd;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int i = 0; i < kHd / 8; ++i) {
        uint4 v;
        v.x = pack_bf16(o[8 * i + 0] * inv, o[8 * i + 1] * inv);
        v.y = pack_bf16(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
        v.z = pack_bf16(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
        v.w = pack_bf16(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
        d4[i] = v;
      }
      p.lse2[(static_cast<size_t>(b) * p.H + h) * p.Lp + row] = m + log2f(l);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int attn_fwd(const void* qkv, void* o, float* lse2, int B, int L, int H, float scale, cudaStream_t st) {
  if (B <= 0 || L <= 0 || H <= 0) return set_error(DCV_ERR_INVALID, "attn_fwd: empty problem");
  const int D = H * kHd;
  CUtensorMap map;
  if (int e = make_tmap_bf16_3d(&map, qkv, (uint64_t)3 * D, (uint64_t)L, (uint64_t)B, (uint64_t)3 * D * 2,
                                (uint64_t)L * 3 * D * 2, kHd, kTq, 1))
    return e;
  static bool attr_done = false;
  if (!attr_done) {
    DCV_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    attr_done = true;
  }
  AttnFwdParams p;
  p.B = B; p.L = L; p.H = H; p.D = D;
  p.Lp = (L + 127) / 128 * 128;
  p.sl2 = scale * 1.4426950408889634f;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.lse2 = lse2;
  dim3 grid((L + kTq - 1) / kTq, H, B);
  ProfScope prof(PT_ATTN_FWD, st);
  attn_fwd_kernel<<<grid, 192, kFwdSmem, st>>>(map, p);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}


// =============================================================================================
// Backward
//   prep   : delta[b,h,q] = sum_d dO[q,d] * O[q,d]                       (HBM-bound)
//   main   : one CTA per (128-key tile, head, image); loops over 128-query tiles:
//              S^T  = K Q^T           dP^T = V dO^T                (SS MMAs, both K-major)
//              P^T  = exp2(S^T*sl2 - lse2[q])     dS^T = P^T o (dP^T - delta[q]) * scale
//              dV  += P^T dO          dK  += dS^T Q                 (A K-major from smem, B MN-major)
//              dQ_i = dS K  (A = the same dS^T smem tile read MN-major, B = K MN-major), added to
//              a fp32 [B,H,L,64] accumulator with one cp.reduce.async.bulk per tile
//   finish : dq_acc fp32 [B,H,L,64] -> bf16 dqkv[:, :, 0:D]
//   TMEM columns: S^T [0,128) | dP^T [128,256) | dV [256,320) | dK [320,384) | dQ [384,448)
// =============================================================================================

struct AttnBwdParams {
  int B, L, H, D, Lp;
  float sl2;    // scale * log2(e)
  float scale;
  const float* lse2;   // [B,H,Lp]
  const float* delta;  // [B,H,Lp]
  float* dq_acc;       // [B,H,L,64] fp32, zero-initialised
  __nv_bfloat16* dqkv; // [B,L,3D]
};

constexpr int kTile16K = 128 * 64 * 2;
constexpr int kBwdSmem = 2 * kTile16K /*K,V*/ + 4 * kTile16K /*Q,dO x2 stages*/ + 2 * 2 * kTile16K /*P,dS*/ +
                         128 * 64 * 4 /*dQ staging*/ + 1024 + 256;

__device__ __forceinline__ void wg_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __launch_bounds__(192, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kTile16K;
  uint8_t* sQ = sV + kTile16K;        // 2 stages
  uint8_t* sdO = sQ + 2 * kTile16K;   // 2 stages
  uint8_t* sP = sdO + 2 * kTile16K;   // [2 q-chunks][128 kv rows][128 B] swizzled
  uint8_t* sdS = sP + 2 * kTile16K;
  float* sdQ = reinterpret_cast<float*>(sdS + 2 * kTile16K);  // [128][64] fp32
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sdQ) + 128 * 64 * 4);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* dp_full = bars + 6;
  uint64_t* p_ready = bars + 7;
  uint64_t* ds_ready = bars + 8;
  uint64_t* dq_full = bars + 9;
  uint64_t* dkv_full = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * kTk;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_q = (p.L + kTq - 1) / kTq;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(ds_ready, 128);
    mbar_init(dq_full, 1);
    mbar_init(dkv_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * kTile16K);
      tma_load_3d(sK, &map_qkv, kv_full, p.D + h * kHd, kv0, b);
      tma_load_3d(sV, &map_qkv, kv_full, 2 * p.D + h * kHd, kv0, b);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        mbar_wait(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * kTile16K);
        tma_load_3d(sQ + st * kTile16K, &map_qkv, &qdo_full[st], h * kHd, i * kTq, b);
        tma_load_3d(sdO + st * kTile16K, &map_do, &qdo_full[st], h * kHd, i * kTq, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);   // S^T, dP^T
      constexpr uint32_t id_kv = make_idesc_bf16(128, 64, 0, 1);   // dV, dK : A K-major, B MN-major
      constexpr uint32_t id_dq = make_idesc_bf16(128, 64, 1, 1);   // dQ     : A MN-major, B MN-major
      const uint64_t dK_k = make_desc_kmajor(smem_u32(sK));
      const uint64_t dV_k = make_desc_kmajor(smem_u32(sV));
      const uint64_t dK_mn = make_desc_mnmajor(smem_u32(sK), kTile16K);
      const uint64_t dP_k0 = make_desc_kmajor(smem_u32(sP));
      const uint64_t dP_k1 = make_desc_kmajor(smem_u32(sP + kTile16K));
      const uint64_t dS_k0 = make_desc_kmajor(smem_u32(sdS));
      const uint64_t dS_k1 = make_desc_kmajor(smem_u32(sdS + kTile16K));
      const uint64_t dS_mn = make_desc_mnmajor(smem_u32(sdS), kTile16K);

      mbar_wait(kv_full, 0);
      mbar_wait(&qdo_full[0], 0);
      tc_fence_after();
      {
        const uint64_t dQ_k = make_desc_kmajor(smem_u32(sQ));
        const uint64_t dO_k = make_desc_kmajor(smem_u32(sdO));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tS, dK_k + 2 * k, dQ_k + 2 * k, id_s, k ? 1u : 0u);
        umma_commit(s_full);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tdP, dV_k + 2 * k, dO_k + 2 * k, id_s, k ? 1u : 0u);
        umma_commit(dp_full);
      }
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        const uint64_t dQ_mn = make_desc_mnmajor(smem_u32(sQ + st * kTile16K), kTile16K);
        const uint64_t dO_mn = make_desc_mnmajor(smem_u32(sdO + st * kTile16K), kTile16K);
        // dV += P^T dO_i
        mbar_wait(p_ready, i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tdV, (k < 4 ? dP_k0 : dP_k1) + 2 * (k & 3), dO_mn + 128 * k, id_kv, (i | k) ? 1u : 0u);
        // S^T of the next query tile may overwrite tS now (phase A of tile i has consumed it)
        uint64_t dQn_k = 0, dOn_k = 0;
        if (i + 1 < n_q) {
          const int st1 = (i + 1) & 1;
          mbar_wait(&qdo_full[st1], ((i + 1) >> 1) & 1);
          tc_fence_after();
          dQn_k = make_desc_kmajor(smem_u32(sQ + st1 * kTile16K));
          dOn_k = make_desc_kmajor(smem_u32(sdO + st1 * kTile16K));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tS, dK_k + 2 * k, dQn_k + 2 * k, id_s, k ? 1u : 0u);
          umma_commit(s_full);
        }
        // dK += dS^T Q_i ; dQ_i = dS K
        mbar_wait(ds_ready, i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tdK, (k < 4 ? dS_k0 : dS_k1) + 2 * (k & 3), dQ_mn + 128 * k, id_kv, (i | k) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ss(tdQ, dS_mn + 128 * k, dK_mn + 128 * k, id_dq, k ? 1u : 0u);
        umma_commit(dq_full);
        umma_commit(&qdo_empty[st]);
        if (i + 1 < n_q) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tdP, dV_k + 2 * k, dOn_k + 2 * k, id_s, k ? 1u : 0u);
          umma_commit(dp_full);
        }
      }
      umma_commit(dkv_full);
    }
  } else {
    // --------------------------- softmax-gradient warpgroup ---------------------------
    const int r = warp * 32 + lane;  // key row (phases A/B) or query row (dQ drain)
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const bool kv_ok = kv0 + r < p.L;
    const size_t stat_base = (static_cast<size_t>(b) * p.H + h) * p.Lp;

    for (int i = 0; i < n_q; ++i) {
      const int q0 = i * kTq;
      const float4* lse4 = reinterpret_cast<const float4*>(p.lse2 + stat_base + q0);
      const float4* del4 = reinterpret_cast<const float4*>(p.delta + stat_base + q0);
      uint32_t pk[64];  // P^T row, packed bf16 (kept for phase B)

      // ---- phase A: P^T = exp2(S^T * sl2 - lse2[q]) ----
      mbar_wait(s_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t sreg[32];
        tmem_ld32(tS + lane_base + c * 32, sreg);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 ls = __ldg(lse4 + c * 8 + g);
          const int qc = q0 + c * 32 + g * 4;
          float p0 = fast_exp2(fmaf(__uint_as_float(sreg[4 * g + 0]), p.sl2, -ls.x));
          float p1 = fast_exp2(fmaf(__uint_as_float(sreg[4 * g + 1]), p.sl2, -ls.y));
          float p2 = fast_exp2(fmaf(__uint_as_float(sreg[4 * g + 2]), p.sl2, -ls.z));
          float p3 = fast_exp2(fmaf(__uint_as_float(sreg[4 * g + 3]), p.sl2, -ls.w));
          if (!kv_ok || qc + 0 >= p.L) p0 = 0.f;
          if (!kv_ok || qc + 1 >= p.L) p1 = 0.f;
          if (!kv_ok || qc + 2 >= p.L) p2 = 0.f;
          if (!kv_ok || qc + 3 >= p.L) p3 = 0.f;
          pk[c * 16 + 2 * g] = pack_bf16(p0, p1);
          pk[c * 16 + 2 * g + 1] = pack_bf16(p2, p3);
        }
        // 32 q-columns = 4 x 16 B into chunk (c/2), 16B-slots (c%2)*4 .. +3 of row r
        uint8_t* base = sP + (c >> 1) * kTile16K;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 val = make_uint4(pk[c * 16 + 4 * v], pk[c * 16 + 4 * v + 1], pk[c * 16 + 4 * v + 2],
                                 pk[c * 16 + 4 * v + 3]);
          *reinterpret_cast<uint4*>(base + sw128_offset(r, (c & 1) * 4 + v)) = val;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);

      // ---- phase B: dS^T = P^T o (dP^T - delta[q]) * scale ----
      mbar_wait(dp_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t dreg[32];
        tmem_ld32(tdP + lane_base + c * 32, dreg);
        tmem_ld_wait();
        uint32_t dk[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 de = __ldg(del4 + c * 8 + g);
          const float2 pa = unpack_bf16(pk[c * 16 + 2 * g]);
          const float2 pb = unpack_bf16(pk[c * 16 + 2 * g + 1]);
          const float d0 = pa.x * (__uint_as_float(dreg[4 * g + 0]) - de.x) * p.scale;
          const float d1 = pa.y * (__uint_as_float(dreg[4 * g + 1]) - de.y) * p.scale;
          const float d2 = pb.x * (__uint_as_float(dreg[4 * g + 2]) - de.z) * p.scale;
          const float d3 = pb.y * (__uint_as_float(dreg[4 * g + 3]) - de.w) * p.scale;
          dk[2 * g] = pack_bf16(d0, d1);
          dk[2 * g + 1] = pack_bf16(d2, d3);
        }
        uint8_t* base = sdS + (c >> 1) * kTile16K;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 val = make_uint4(dk[4 * v], dk[4 * v + 1], dk[4 * v + 2], dk[4 * v + 3]);
          *reinterpret_cast<uint4*>(base + sw128_offset(r, (c & 1) * 4 + v)) = val;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_ready);

      // ---- drain dQ_i: TMEM -> smem (fp32) -> one bulk reduce-add into dq_acc ----
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      if (r == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous reduce has read sdQ
      wg_barrier();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t qreg[32];
        tmem_ld32(tdQ + lane_base + c * 32, qreg);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(sdQ + r * 64 + c * 32);
#pragma unroll
        for (int v = 0; v < 8; ++v)
          dst[v] = make_float4(__uint_as_float(qreg[4 * v]), __uint_as_float(qreg[4 * v + 1]),
                               __uint_as_float(qreg[4 * v + 2]), __uint_as_float(qreg[4 * v + 3]));
      }
      tc_fence_before();
      fence_proxy_async_smem();
      wg_barrier();
      if (r == 0) {
        const int rows = min(kTq, p.L - q0);
        float* gdst = p.dq_acc + ((static_cast<size_t>(b) * p.H + h) * p.L + q0) * kHd;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
                     "r"(smem_u32(sdQ)), "r"(rows * kHd * 4)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }

    // ---- epilogue: dK, dV rows of this key tile ----
    mbar_wait(dkv_full, 0);
    tc_fence_after();
#pragma unroll
    for (int which = 0; which < 2; ++which) {  // 0: dK -> column block D, 1: dV -> column block 2D
      const uint32_t tsrc = which == 0 ? tdK : tdV;
      __nv_bfloat16* dst = p.dqkv + (static_cast<size_t>(b) * p.L + kv0 + r) * (3 * p.D) + (which + 1) * p.D + h * kHd;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t a[32];
        tmem_ld32(tsrc + lane_base + c * 32, a);
        tmem_ld_wait();
        if (kv_ok) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(a[8 * v + 0]), __uint_as_float(a[8 * v + 1]));
            o.y = pack_bf16(__uint_as_float(a[8 * v + 2]), __uint_as_float(a[8 * v + 3]));
            o.z = pack_bf16(__uint_as_float(a[8 * v + 4]), __uint_as_float(a[8 * v + 5]));
            o.w = pack_bf16(__uint_as_float(a[8 * v + 6]), __uint_as_float(a[8 * v + 7]));
            d4[v] = o;
          }
        }
      }
    }
    if (r == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // reduces fully performed
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[b,h,q] = sum_d dO*O ; 8 threads per (token, head), uint4 (8 x bf16) each
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dO,
                                     float* __restrict__ delta, int B, int L, int H, int Lp) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * L * H * 8;
  const bool ok = gid < total;
  float acc = 0.f;
  long long grp = gid >> 3;  // (token, head)
  if (ok) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(o) + gid);
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(dO) + gid);
    const uint32_t* au = reinterpret_cast<const uint32_t*>(&a);
    const uint32_t* gu = reinterpret_cast<const uint32_t*>(&g);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_bf16(au[i]), y = unpack_bf16(gu[i]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (ok && (gid & 7) == 0) {
    const int hh = static_cast<int>(grp % H);
    const long long tok = grp / H;
    const int q = static_cast<int>(tok % L);
    const int bb = static_cast<int>(tok / L);
    delta[(static_cast<size_t>(bb) * H + hh) * Lp + q] = acc;
  }
}

// dq_acc fp32 [B,H,L,64] -> dqkv bf16 [B,L,3D] columns [h*64, h*64+64)
__global__ void attn_bwd_finish_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, int B,
                                       int L, int H) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // one per 8 elements
  const long long total = static_cast<long long>(B) * H * L * 8;
  if (gid >= total) return;
  const int c8 = static_cast<int>(gid & 7);
  long long t = gid >> 3;
  const int q = static_cast<int>(t % L);
  t /= L;
  const int hh = static_cast<int>(t % H);
  const int bb = static_cast<int>(t / H);
  const float4* src = reinterpret_cast<const float4*>(dq_acc) + gid * 2;
  const float4 a = __ldg(src), c = __ldg(src + 1);
  uint4 o;
  o.x = pack_bf16(a.x, a.y);
  o.y = pack_bf16(a.z, a.w);
  o.z = pack_bf16(c.x, c.y);
  o.w = pack_bf16(c.z, c.w);
  const int D = H * kHd;
  *reinterpret_cast<uint4*>(dqkv + (static_cast<size_t>(bb) * L + q) * (3 * D) + hh * kHd + c8 * 8) = o;
}

int attn_bwd(const void* qkv, const void* o, const void* dO, const float* lse2, float* delta, float* dq_acc,
             void* dqkv, int B, int L, int H, float scale, cudaStream_t st) {
  if (B <= 0 || L <= 0 || H <= 0) return set_error(DCV_ERR_INVALID, "attn_bwd: empty problem");
  const int D = H * kHd;
  const int Lp = (L + 127) / 128 * 128;
  CUtensorMap map_qkv, map_do;
  if (int e = make_tmap_bf16_3d(&map_qkv, qkv, (uint64_t)3 * D, (uint64_t)L, (uint64_t)B, (uint64_t)3 * D * 2,
                                (uint64_t)L * 3 * D * 2, kHd, kTq, 1))
    return e;
  if (int e = make_tmap_bf16_3d(&map_do, dO, (uint64_t)D, (uint64_t)L, (uint64_t)B, (uint64_t)D * 2,
                                (uint64_t)L * D * 2, kHd, kTq, 1))
    return e;
  static bool attr_done = false;
  if (!attr_done) {
    DCV_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    attr_done = true;
  }
  {
    ProfScope prof(PT_ATTN_BWD_PREP, st);
    const long long total = static_cast<long long>(B) * L * H * 8;
    attn_bwd_prep_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(dO), delta, B, L, H, Lp);
    DCV_CUDA(cudaGetLastError());
    DCV_CUDA(cudaMemsetAsync(dq_acc, 0, static_cast<size_t>(B) * H * L * kHd * sizeof(float), st));
  }
  AttnBwdParams p;
  p.B = B; p.L = L; p.H = H; p.D = D; p.Lp = Lp;
  p.scale = scale;
  p.sl2 = scale * 1.4426950408889634f;
  p.lse2 = lse2; p.delta = delta; p.dq_acc = dq_acc;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  dim3 grid((L + kTk - 1) / kTk, H, B);
  {
    ProfScope prof(PT_ATTN_BWD, st);
    attn_bwd_kernel<<<grid, 192, kBwdSmem, st>>>(map_qkv, map_do, p);
    DCV_CUDA(cudaGetLastError());
  }
  {
    ProfScope prof(PT_ATTN_BWD_FIN, st);
    const long long total = static_cast<long long>(B) * H * L * 8;
    attn_bwd_finish_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
        dq_acc, reinterpret_cast<__nv_bfloat16*>(dqkv), B, L, H);
    DCV_CUDA(cudaGetLastError());
  }
  count_launch(3);
  return 0;
}

}  // namespace dcv
