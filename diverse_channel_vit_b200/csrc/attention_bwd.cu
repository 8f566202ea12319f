// Flash-attention backward on tcgen05 / TMEM (head_dim 64), replacing the autograd backward of
// reference models/vit.py:126-141 (softmax(q k^T * scale) v).
//
//   prep   : delta[b,h,q] = sum_d dO[q,d] * O[q,d] (0 on the pad rows), dq accumulator cleared
//   main   : one CTA per (128-key tile, head, image), 1 CTA / SM, looping over 128-query tiles:
//              S^T  = K Q^T            dP^T = V dO^T                    (SS MMAs, operands K-major)
//              P^T  = exp2(S^T*sl2 - lse2[q])          -> bf16 into TMEM (A operand of the dV MMA)
//              dS^T = P^T o (dP^T - delta[q])          -> bf16 into smem (hand-swizzled UMMA tile)
//              dV  += P^T dO   (TS MMA)     dK += dS^T Q   (SS)      dQ_i = dS K   (SS, A MN-major)
//            dQ_i is drained TMEM -> swizzled smem -> one cp.reduce.async.bulk.tensor (fp32 add) per
//            64x... half tile into the [B,H,L,64] accumulator; the softmax scale is folded into the
//            dK epilogue and the finish kernel.
//   finish : scale * dq_acc fp32 [B,H,L,64] -> bf16 dqkv[:, :, 0:D]
//
// Warp roles (320 threads): warps 0-7 = two compute warpgroups, warpgroup g owns the query columns
// [64g, 64g+64) of every S^T / dP^T tile (TMEM lane = key row = 32*(warp%4)+lane) and the dQ columns
// [32g, 32g+32); warp 8 = TMA producer; warp 9 = MMA issuer + TMEM owner.
// TMEM columns: S^T [0,128) | dP^T [128,256) | dV [256,320) | dK [320,384) | dQ [384,448) | P^T bf16 [448,512)
#include "common.cuh"
#include "host.h"

namespace dcv {

namespace {

constexpr int kHd = 64;
constexpr int kTq = 128;
constexpr int kTk = 128;
constexpr int kTile16K = 128 * 64 * 2;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// debug timeline: when non-null, CTA (1,0,0) records clock64() stamps: slot = role * 1024 + iter * 8 + point
__device__ long long* g_attn_timeline = nullptr;
#define TL(role, it, pt)                                                                      \
  do {                                                                                        \
    if (tl) tl[(role) * 1024 + (it) * 8 + (pt)] = clock64();                                   \
  } while (0)

struct AttnBwdParams {
  int B, L, H, D, Lp;
  int n_q;      // query tiles to visit (all, or 1 when only the CLS rows carry gradient)
  float sl2;    // scale * log2(e)
  float scale;
  const float* lse2;   // [B,H,Lp]
  const float* delta;  // [B,H,Lp]
  __nv_bfloat16* dqkv; // [B,L,3D]
};

// smem: K,V | Q,dO x2 stages | dS^T (2 chunks of [128 kv][64 q]) | dQ staging (2 boxes of [128 q][32] fp32)
//       | lse2 / delta of the query tile x2 stages (TMA bulk copies riding on the Q/dO barrier)
constexpr int kStatBytes = 2 * kTq * 4;  // 128 lse2 + 128 delta
constexpr int kBwdSmem = 2 * kTile16K + 4 * kTile16K + 2 * kTile16K + 2 * kTile16K + 2 * kStatBytes + 1024 + 256;
constexpr int kBwdThreads = 448;  // 8 compute warps + TMA warp + MMA warp + 4 dQ-drain warps

__device__ __forceinline__ void wg_barrier(int g) {  // named barrier 1 + g, the 128 threads of warpgroup g
  asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
}

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                const __grid_constant__ CUtensorMap map_dq, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kTile16K;
  uint8_t* sQ = sV + kTile16K;        // 2 stages
  uint8_t* sdO = sQ + 2 * kTile16K;   // 2 stages
  uint8_t* sdS = sdO + 2 * kTile16K;  // [2 q-chunks][128 kv rows][128 B] swizzled
  uint8_t* sdQ = sdS + 2 * kTile16K;  // [2 d-halves][128 q rows][32 fp32] swizzled
  uint8_t* sStat = sdQ + 2 * kTile16K;  // [2 stages][lse2 128 | delta 128] fp32
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 2 * kStatBytes);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* dp_full = bars + 6;
  uint64_t* p_ready = bars + 7;
  uint64_t* ds_ready = bars + 8;
  uint64_t* dq_full = bars + 9;
  uint64_t* dq_empty = bars + 10;
  uint64_t* dkv_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * kTk;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_q = p.n_q;
  long long* tl = (g_attn_timeline && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 &&
                   (warp == 9 || warp == 0 || warp == 4))
                      ? g_attn_timeline
                      : nullptr;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_dq);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(ds_ready, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 128);  // the four drain warps
    mbar_init(dkv_full, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384, tP = tmem_base + 448;

  if (warp == 8) {
    // ------------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * kTile16K);
      tma_load_3d(sK, &map_qkv, kv_full, p.D + h * kHd, kv0, b);
      tma_load_3d(sV, &map_qkv, kv_full, 2 * p.D + h * kHd, kv0, b);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        mbar_wait(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * kTile16K + kStatBytes);
        tma_load_3d(sQ + st * kTile16K, &map_qkv, &qdo_full[st], h * kHd, i * kTq, b);
        tma_load_3d(sdO + st * kTile16K, &map_do, &qdo_full[st], h * kHd, i * kTq, b);
        const size_t so = (static_cast<size_t>(b) * p.H + h) * p.Lp + static_cast<size_t>(i) * kTq;
        bulk_load_1d(sStat + st * kStatBytes, p.lse2 + so, kTq * 4, &qdo_full[st]);
        bulk_load_1d(sStat + st * kStatBytes + kTq * 4, p.delta + so, kTq * 4, &qdo_full[st]);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------- MMA issuer -------------------------------------
    // The whole warp walks the loop (all lanes wait on the barriers) and ONE elected lane issues the MMAs /
    // commits inside warp-uniform control flow: under `if (lane == 0)` the compiler has to assume divergent
    // operands and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop, which made the
    // single issuing thread -- not the tensor pipe -- the bottleneck of this kernel.
    {
      constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);   // S^T, dP^T
      constexpr uint32_t id_kv = make_idesc_bf16(128, 64, 0, 1);   // dV, dK : A (TMEM) K-major, B MN-major
      constexpr uint32_t id_dq = make_idesc_bf16(128, 64, 1, 1);   // dQ     : A MN-major, B MN-major
      const uint64_t dK_k = make_desc_kmajor(smem_u32(sK));
      const uint64_t dV_k = make_desc_kmajor(smem_u32(sV));
      const uint64_t dK_mn = make_desc_mnmajor(smem_u32(sK), kTile16K);
      const uint64_t dS_mn = make_desc_mnmajor(smem_u32(sdS), kTile16K);
      const uint64_t dS_k0 = make_desc_kmajor(smem_u32(sdS));
      const uint64_t dS_k1 = make_desc_kmajor(smem_u32(sdS + kTile16K));

      mbar_wait(kv_full, 0);
      mbar_wait(&qdo_full[0], 0);
      tc_fence_after();
      {
        const uint64_t dQ_k = make_desc_kmajor(smem_u32(sQ));
        const uint64_t dO_k = make_desc_kmajor(smem_u32(sdO));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tS, dK_k + 2 * k, dQ_k + 2 * k, id_s, k ? 1u : 0u);
          umma_commit(s_full);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tdP, dV_k + 2 * k, dO_k + 2 * k, id_s, k ? 1u : 0u);
          umma_commit(dp_full);
        }
        __syncwarp();
      }
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        const uint64_t dQ_mn = make_desc_mnmajor(smem_u32(sQ + st * kTile16K), kTile16K);
        const uint64_t dO_mn = make_desc_mnmajor(smem_u32(sdO + st * kTile16K), kTile16K);
        const uint32_t acc = i ? 1u : 0u;
        // dV += P^T dO_i  (A = P^T from TMEM: 16 q per K step = 8 packed columns)
        TL(0, i, 0);
        mbar_wait(p_ready, i & 1);
        tc_fence_after();
        TL(0, i, 1);
        if (elect_one()) {
          umma_ts(tdV, tP, dO_mn, id_kv, acc);
#pragma unroll
          for (int k = 1; k < 8; ++k) umma_ts(tdV, tP + 8 * k, dO_mn + 128 * k, id_kv, 1u);
        }
        __syncwarp();
        // S^T of the next query tile may overwrite tS now (phase A of tile i has consumed it)
        uint64_t dOn_k = 0;
        if (i + 1 < n_q) {
          const int st1 = (i + 1) & 1;
          mbar_wait(&qdo_full[st1], ((i + 1) >> 1) & 1);
          tc_fence_after();
          const uint64_t dQn_k = make_desc_kmajor(smem_u32(sQ + st1 * kTile16K));
          dOn_k = make_desc_kmajor(smem_u32(sdO + st1 * kTile16K));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tS, dK_k + 2 * k, dQn_k + 2 * k, id_s, k ? 1u : 0u);
            umma_commit(s_full);
          }
          __syncwarp();
        }
        // dK += dS^T Q_i ; dQ_i = dS K
        TL(0, i, 2);
        mbar_wait(ds_ready, i & 1);
        tc_fence_after();
        TL(0, i, 3);
        if (elect_one()) {
          umma_ss(tdK, dS_k0, dQ_mn, id_kv, acc);
#pragma unroll
          for (int k = 1; k < 8; ++k)
            umma_ss(tdK, (k < 4 ? dS_k0 : dS_k1) + 2 * (k & 3), dQ_mn + 128 * k, id_kv, 1u);
        }
        __syncwarp();
        if (i > 0) {  // the compute warpgroups have drained dQ_{i-1} out of TMEM
          mbar_wait(dq_empty, (i - 1) & 1);
          tc_fence_after();
        }
        TL(0, i, 4);
        if (elect_one()) {
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
        __syncwarp();
        TL(0, i, 5);
      }
      if (elect_one()) umma_commit(dkv_full);
      __syncwarp();
    }
  } else if (warp >= 10) {
    // ------------------------------------ dQ drain warps ------------------------------------
    // dQ_i (128 queries x 64) : TMEM -> two swizzled [128][32] fp32 boxes in smem -> TMA reduce-add into the
    // [B,H,L,64] accumulator.  Off the critical path of the compute warpgroups.
    const int q4 = warp & 3;  // TMEM lane quadrant (warps 10..13 -> quadrants 2,3,0,1)
    const int r = q4 * 32 + lane;
    const int tid_d = threadIdx.x - 320;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    const uint32_t sdQ0 = smem_u32(sdQ), sdQ1 = smem_u32(sdQ + kTile16K);
    for (int i = 0; i < n_q; ++i) {
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t q0r[32], q1r[32];
      tmem_ld32(tdQ + lane_base, q0r);
      tmem_ld32(tdQ + lane_base + 32, q1r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_empty);
      if (tid_d == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous reduce has read the boxes
      asm volatile("bar.sync 3, 128;" ::: "memory");
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        st_shared_v4(sdQ0 + sw128_offset(r, v), q0r[4 * v], q0r[4 * v + 1], q0r[4 * v + 2], q0r[4 * v + 3]);
        st_shared_v4(sdQ1 + sw128_offset(r, v), q1r[4 * v], q1r[4 * v + 1], q1r[4 * v + 2], q1r[4 * v + 3]);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 3, 128;" ::: "memory");
      if (tid_d == 0) {
        tma_reduce_add_3d(&map_dq, sdQ, 0, i * kTq, b * p.H + h);
        tma_reduce_add_3d(&map_dq, sdQ + kTile16K, 32, i * kTq, b * p.H + h);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (tid_d == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // reduces fully performed
  } else {
    // --------------------------------- compute warpgroups ---------------------------------
    const int g = warp >> 2;                  // column half
    const int q4 = warp & 3;                  // TMEM lane quadrant
    const int r = q4 * 32 + lane;             // key row (phases A/B) or query row (dQ drain)
    const int tid_g = threadIdx.x & 127;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    const bool kv_ok = kv0 + r < p.L;
    const bool kv_tail = kv0 + kTk > p.L;     // uniform: only the last key tile has masked rows
    const uint32_t sdS_g = smem_u32(sdS + g * kTile16K);
    const uint32_t sdQ_g = smem_u32(sdQ + g * kTile16K);

    for (int i = 0; i < n_q; ++i) {
      // lse2 / delta of this warpgroup's 64 queries (smem, broadcast reads)
      const uint32_t s_lse = smem_u32(sStat + (i & 1) * kStatBytes) + g * 256;
      const uint32_t s_del = s_lse + kTq * 4;
      float pf[64];  // P^T row (64 queries) in fp32, kept for phase B

      // ---- phase A: P^T = exp2(S^T * sl2 - lse2[q]) ----
      TL(1 + g, i, 0);
      mbar_wait(&qdo_full[i & 1], (i >> 1) & 1);  // stats landed (completes long before S_i)
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      TL(1 + g, i, 1);
      {
        uint32_t s0[32], s1[32];
        tmem_ld32(tS + lane_base + g * 64, s0);
        tmem_ld32(tS + lane_base + g * 64 + 32, s1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 la = ld_shared_f4(s_lse + j * 16), lb = ld_shared_f4(s_lse + 128 + j * 16);
          pf[4 * j + 0] = fast_exp2(fmaf(__uint_as_float(s0[4 * j + 0]), p.sl2, -la.x));
          pf[4 * j + 1] = fast_exp2(fmaf(__uint_as_float(s0[4 * j + 1]), p.sl2, -la.y));
          pf[4 * j + 2] = fast_exp2(fmaf(__uint_as_float(s0[4 * j + 2]), p.sl2, -la.z));
          pf[4 * j + 3] = fast_exp2(fmaf(__uint_as_float(s0[4 * j + 3]), p.sl2, -la.w));
          pf[32 + 4 * j + 0] = fast_exp2(fmaf(__uint_as_float(s1[4 * j + 0]), p.sl2, -lb.x));
          pf[32 + 4 * j + 1] = fast_exp2(fmaf(__uint_as_float(s1[4 * j + 1]), p.sl2, -lb.y));
          pf[32 + 4 * j + 2] = fast_exp2(fmaf(__uint_as_float(s1[4 * j + 2]), p.sl2, -lb.z));
          pf[32 + 4 * j + 3] = fast_exp2(fmaf(__uint_as_float(s1[4 * j + 3]), p.sl2, -lb.w));
        }
      }
      if (kv_tail && !kv_ok) {
#pragma unroll
        for (int j = 0; j < 64; ++j) pf[j] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(pf[c * 32 + 2 * j], pf[c * 32 + 2 * j + 1]);
        tmem_st16(tP + lane_base + g * 32 + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
      TL(1 + g, i, 2);

      // ---- phase B: dS^T = P^T o (dP^T - delta[q])   (softmax scale folded into dK / dQ epilogues) ----
      TL(1 + g, i, 3);
      mbar_wait(dp_full, i & 1);
      tc_fence_after();
      TL(1 + g, i, 4);
      {
        uint32_t d0[32], d1[32];
        tmem_ld32(tdP + lane_base + g * 64, d0);
        tmem_ld32(tdP + lane_base + g * 64 + 32, d1);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 8; ++v) {  // 8 queries -> one 16-byte slot of the swizzled dS^T tile
          const uint32_t* dr = v < 4 ? d0 : d1;
          const int o = (v & 3) * 8;
          const float4 da = ld_shared_f4(s_del + v * 32), db = ld_shared_f4(s_del + v * 32 + 16);
          const float e0 = pf[v * 8 + 0] * (__uint_as_float(dr[o + 0]) - da.x);
          const float e1 = pf[v * 8 + 1] * (__uint_as_float(dr[o + 1]) - da.y);
          const float e2 = pf[v * 8 + 2] * (__uint_as_float(dr[o + 2]) - da.z);
          const float e3 = pf[v * 8 + 3] * (__uint_as_float(dr[o + 3]) - da.w);
          const float e4 = pf[v * 8 + 4] * (__uint_as_float(dr[o + 4]) - db.x);
          const float e5 = pf[v * 8 + 5] * (__uint_as_float(dr[o + 5]) - db.y);
          const float e6 = pf[v * 8 + 6] * (__uint_as_float(dr[o + 6]) - db.z);
          const float e7 = pf[v * 8 + 7] * (__uint_as_float(dr[o + 7]) - db.w);
          st_shared_v4(sdS_g + sw128_offset(r, v), pack_bf16(e0, e1), pack_bf16(e2, e3), pack_bf16(e4, e5),
                       pack_bf16(e6, e7));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_ready);
      TL(1 + g, i, 5);
    }
    // ---- epilogue: dK (x scale) and dV rows of this key tile; warpgroup g writes d columns [32g, 32g+32) ----
    mbar_wait(dkv_full, 0);
    tc_fence_after();
#pragma unroll
    for (int which = 0; which < 2; ++which) {  // 0: dK -> column block D, 1: dV -> column block 2D
      const uint32_t tsrc = which == 0 ? tdK : tdV;
      const float mul = which == 0 ? p.scale : 1.0f;
      uint32_t a[32];
      tmem_ld32(tsrc + lane_base + g * 32, a);
      tmem_ld_wait();
      if (kv_ok) {
        __nv_bfloat16* dst = p.dqkv + (static_cast<size_t>(b) * p.L + kv0 + r) * (3 * p.D) + (which + 1) * p.D +
                             h * kHd + g * 32;
        uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(a[8 * v + 0]) * mul, __uint_as_float(a[8 * v + 1]) * mul);
          o.y = pack_bf16(__uint_as_float(a[8 * v + 2]) * mul, __uint_as_float(a[8 * v + 3]) * mul);
          o.z = pack_bf16(__uint_as_float(a[8 * v + 4]) * mul, __uint_as_float(a[8 * v + 5]) * mul);
          o.w = pack_bf16(__uint_as_float(a[8 * v + 6]) * mul, __uint_as_float(a[8 * v + 7]) * mul);
          d4[v] = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[b,h,q] = sum_d dO*O ; 8 threads per (token, head), uint4 (8 x bf16) each; pad rows [L, Lp) <- 0
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
  // pad rows
  const int pad = Lp - L;
  if (gid < static_cast<long long>(B) * H * pad) {
    const int j = static_cast<int>(gid % pad);
    const long long bh = gid / pad;
    delta[bh * Lp + L + j] = 0.f;
  }
}

// CLS-only variant: dO compact [B, D]; delta[b,h,0] = dO[b] . o[b, row 0]; delta[b,h,1..127] = 0
__global__ void attn_bwd_prep_cls_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dO,
                                         float* __restrict__ delta, int B, int L, int H, int Lp) {
  const int bh = blockIdx.x;  // one warp-sized CTA per (image, head)
  const int b = bh / H, h = bh - b * H;
  const int D = H * kHd;
  const int lane = threadIdx.x;
  const __nv_bfloat162 a = reinterpret_cast<const __nv_bfloat162*>(o + static_cast<size_t>(b) * L * D + h * kHd)[lane];
  const __nv_bfloat162 g = reinterpret_cast<const __nv_bfloat162*>(dO + static_cast<size_t>(b) * D + h * kHd)[lane];
  const float2 x = __bfloat1622float2(a), y = __bfloat1622float2(g);
  const float acc = warp_sum(x.x * y.x + x.y * y.y);
  float* dst = delta + static_cast<size_t>(bh) * Lp;
  for (int i = lane; i < 128; i += 32) dst[i] = (i == 0) ? acc : 0.f;
}

// scale * dq_acc fp32 [B,H,L,64] -> dqkv bf16 [B,L,3D] columns [h*64, h*64+64)
__global__ void attn_bwd_finish_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, int B,
                                       int L, int H, float scale) {
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
  o.x = pack_bf16(a.x * scale, a.y * scale);
  o.y = pack_bf16(a.z * scale, a.w * scale);
  o.z = pack_bf16(c.x * scale, c.y * scale);
  o.w = pack_bf16(c.z * scale, c.w * scale);
  const int D = H * kHd;
  *reinterpret_cast<uint4*>(dqkv + (static_cast<size_t>(bb) * L + q) * (3 * D) + hh * kHd + c8 * 8) = o;
}

}  // namespace

int debug_attn_timeline(long long* buf) {
  DCV_CUDA(cudaMemcpyToSymbol(g_attn_timeline, &buf, sizeof(buf)));
  return 0;
}

int attn_bwd(const void* qkv, const void* o, const void* dO, const float* lse2, float* delta, float* dq_acc,
             void* dqkv, int B, int L, int H, float scale, cudaStream_t st, bool cls_only) {
  if (B <= 0 || L <= 0 || H <= 0) return set_error(DCV_ERR_INVALID, "attn_bwd: empty problem");
  const int D = H * kHd;
  const int Lp = (L + 127) / 128 * 128;
  CUtensorMap map_qkv, map_do, map_dq;
  if (int e = make_tmap_bf16_3d(&map_qkv, qkv, (uint64_t)3 * D, (uint64_t)L, (uint64_t)B, (uint64_t)3 * D * 2,
                                (uint64_t)L * 3 * D * 2, kHd, kTq, 1))
    return e;
  // cls_only: one gradient row per image; rows 1..127 of the query tile come back as TMA zero fill
  const uint64_t do_rows = cls_only ? 1 : (uint64_t)L;
  if (int e = make_tmap_bf16_3d(&map_do, dO, (uint64_t)D, do_rows, (uint64_t)B, (uint64_t)D * 2,
                                do_rows * D * 2, kHd, kTq, 1))
    return e;
  if (int e = make_tmap_f32_3d(&map_dq, dq_acc, 64, (uint64_t)L, (uint64_t)B * H, 64 * 4, (uint64_t)L * 64 * 4, 32, kTq, 1))
    return e;
  static bool attr_done = false;
  if (!attr_done) {
    DCV_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    attr_done = true;
  }
  {
    ProfScope prof(PT_ATTN_BWD_PREP, st);
    if (cls_only) {
      attn_bwd_prep_cls_kernel<<<B * H, 32, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(o),
                                                     reinterpret_cast<const __nv_bfloat16*>(dO), delta, B, L, H, Lp);
    } else {
      const long long total = static_cast<long long>(B) * L * H * 8;
      attn_bwd_prep_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
          reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(dO), delta, B, L, H, Lp);
    }
    DCV_CUDA(cudaGetLastError());
    DCV_CUDA(cudaMemsetAsync(dq_acc, 0, static_cast<size_t>(B) * H * L * kHd * sizeof(float), st));
  }
  AttnBwdParams p;
  p.B = B; p.L = L; p.H = H; p.D = D; p.Lp = Lp;
  p.n_q = cls_only ? 1 : (L + kTq - 1) / kTq;
  p.scale = scale;
  p.sl2 = scale * 1.4426950408889634f;
  p.lse2 = lse2; p.delta = delta;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  dim3 grid((L + kTk - 1) / kTk, H, B);
  {
    ProfScope prof(PT_ATTN_BWD, st);
    attn_bwd_kernel<<<grid, kBwdThreads, kBwdSmem, st>>>(map_qkv, map_do, map_dq, p);
    DCV_CUDA(cudaGetLastError());
  }
  {
    ProfScope prof(PT_ATTN_BWD_FIN, st);
    const long long total = static_cast<long long>(B) * H * L * 8;
    attn_bwd_finish_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
        dq_acc, reinterpret_cast<__nv_bfloat16*>(dqkv), B, L, H, scale);
    DCV_CUDA(cudaGetLastError());
  }
  count_launch(3);
  return 0;
}

}  // namespace dcv
