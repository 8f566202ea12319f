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

// debug timeline (-DDCV_ATTN_TIMELINE): CTA (1,0,0) records clock64() stamps: slot = role * 1024 + iter * 8 + point
__device__ long long* g_fwd_timeline = nullptr;
#ifdef DCV_ATTN_TIMELINE
#define TLF(role, it, pt)                                                  \
  do {                                                                     \
    if (tl) tl[(role) * 1024 + (it) * 8 + (pt)] = clock64();                \
  } while (0)
#else
#define TLF(role, it, pt) ((void)tl)
#endif
static int g_fwd_mode = 2;
void debug_set_attn_fwd_mode(int m) { g_fwd_mode = m; }
int debug_fwd_timeline(long long* buf) {
  DCV_CUDA(cudaMemcpyToSymbol(g_fwd_timeline, &buf, sizeof(buf)));
  return 0;
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

struct FwdBars {
  uint64_t *s_full, *s_consumed, *p_full, *p_free;
};

// One KV tile of a softmax thread (query row = TMEM lane): NC = number of 32-key chunks of the tile that hold keys --
// 4 for every tile but a ragged last one, whose S MMA only covers ceil16(valid keys) columns and whose PV MMA only
// those K steps.  A compile-time NC keeps the hot NC = 4 instance one straight-line block (a run-time chunk guard
// inside the unrolled body cost 14 % of the kernel: the scheduler no longer interleaved the chunks).
template <int NC, bool PK, int PM>
__device__ __forceinline__ void fwd_softmax_tile(int j, int L, const FwdBars& bars, uint32_t tS, uint32_t tP, uint32_t tO,
                                                 uint32_t lane_base, float sl2, float& m, float& l, long long* tl) {
  const int kv0 = j * kTk;
  const bool tail = kv0 + kTk > L;
  TLF(1, j, 0);
  mbar_wait(bars.s_full, j & 1);
  tc_fence_after();
  TLF(1, j, 1);
  uint32_t sr[32 * NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) tmem_ld32(tS + lane_base + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&sr[32 * c]));
  tmem_ld_wait();
  tc_fence_before();
  mbar_arrive(bars.s_consumed);  // S_{j+1} may overwrite tS
  TLF(1, j, 2);
  if (tail) {
#pragma unroll
    for (int i = 0; i < 32 * NC; ++i)
      if (kv0 + i >= L) sr[i] = 0xff800000u;  // -inf
  }
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
  if constexpr (PK) {
#pragma unroll
    for (int i = 0; i < 32 * NC; i += 8) {
      mx0 = fmax3(mx0, __uint_as_float(sr[i]), __uint_as_float(sr[i + 1]));
      mx1 = fmax3(mx1, __uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3]));
      mx2 = fmax3(mx2, __uint_as_float(sr[i + 4]), __uint_as_float(sr[i + 5]));
      mx3 = fmax3(mx3, __uint_as_float(sr[i + 6]), __uint_as_float(sr[i + 7]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32 * NC; i += 4) {
      mx0 = fmaxf(mx0, __uint_as_float(sr[i]));
      mx1 = fmaxf(mx1, __uint_as_float(sr[i + 1]));
      mx2 = fmaxf(mx2, __uint_as_float(sr[i + 2]));
      mx3 = fmaxf(mx3, __uint_as_float(sr[i + 3]));
    }
  }
  const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;
  const bool grow = mx > m + 8.0f;  // first tile: m = -inf -> true
  bool pv_done = j == 0;            // has this thread observed the completion of PV_{j-1} ?
  if (__any_sync(0xffffffffu, grow)) {
    const float alpha = grow ? fast_exp2(m - mx) : 1.0f;  // exp2(-inf) = 0 on the first tile
    if (grow) {
      m = mx;
      l *= alpha;
    }
    if (j > 0) {  // rescale the O accumulator in TMEM (warp-collective; lanes that did not grow use 1)
      mbar_wait(bars.p_free, (j - 1) & 1);
      tc_fence_after();
      pv_done = true;
#pragma unroll
      for (int c = 0; c < kHd / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(tO + lane_base + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st16(tO + lane_base + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
        tmem_st16(tO + lane_base + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&o[16]));
      }
    }
  }
  TLF(1, j, 3);
  // exponentials in place (S_{j+1} and PV_{j-1} run on the tensor pipe meanwhile)
  if constexpr (PK) {
    const uint64_t sl2x2 = pack_f32x2(sl2, sl2), negm = pack_f32x2(-m, -m);
    uint64_t sum_a = pack_f32x2(0.f, 0.f), sum_b = sum_a;
#pragma unroll
    for (int i = 0; i < 16 * NC; ++i) {
      const uint64_t x = fma_f32x2(pack_f32x2(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), sl2x2, negm);
      uint64_t pr;
      float p0, p1;
      if ((PM >> (i & 7)) & 1) {
        pr = exp2_poly_f32x2(x);
        unpack_f32x2(pr, p0, p1);
      } else {
        unpack_f32x2(x, p0, p1);
        p0 = fast_exp2(p0);
        p1 = fast_exp2(p1);
        pr = pack_f32x2(p0, p1);
      }
      if (i & 1) sum_b = add_f32x2(sum_b, pr); else sum_a = add_f32x2(sum_a, pr);
      sr[i] = pack_bf16(p0, p1);
    }
    float s0, s1;
    unpack_f32x2(add_f32x2(sum_a, sum_b), s0, s1);
    l += s0 + s1;
  } else {
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32 * NC; i += 2) {
      const float p0 = fast_exp2(fmaf(__uint_as_float(sr[i]), sl2, -m));
      const float p1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), sl2, -m));
      sum0 += p0;
      sum1 += p1;
      sr[i >> 1] = pack_bf16(p0, p1);
    }
    l += sum0 + sum1;
  }
  TLF(1, j, 4);
  if (!pv_done) {  // PV_{j-1} has finished reading P_{j-1}
    mbar_wait(bars.p_free, (j - 1) & 1);
    tc_fence_after();
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) tmem_st16(tP + lane_base + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&sr[c * 16]));
  tmem_st_wait();
  tc_fence_before();
  mbar_arrive(bars.p_full);
  TLF(1, j, 5);
}

// Pipeline inside one CTA (a second CTA on the same SM fills the MUFU while this one waits):
//   softmax j : read S_j into registers -> [s_consumed] -> row max -> exponentials in place -> P_j to TMEM -> [p_full]
//   MMA warp  : S_{j+1} = Q K_{j+1}^T is issued at s_consumed(j), i.e. it runs underneath the exponentials of
//               tile j; PV_j at p_full(j).  K and V ride separate 2-stage rings: the K slot is free again as soon
//               as S_j has completed, the V slot after PV_j.
// PK: packed-pair fp32 math (FFMA2 / FADD2) and 3-input max in the softmax; PM: 8-bit mask over the pairs of every group
// of 8 score pairs whose exponentials are evaluated on the FMA pipe (exp2_poly_f32x2) instead of the MUFU.
template <bool PK, int PM>
__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTq * kHd * 2;                       // kFwdStages x 16 KB
  uint8_t* sV = sK + kFwdStages * kTk * kHd * 2;            // kFwdStages x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kFwdStages * kTk * kHd * 2);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* s_consumed = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* p_free = bars + 12;   // PV_j complete: P and O are quiescent
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTq;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.L + kTk - 1) / kTk;
  const int kv_valid_tail = p.L - (n_kv - 1) * kTk;   // keys in the last KV tile (1..128)
  const int nt16 = (kv_valid_tail + 15) & ~15;        // ... rounded up to the MMA's N / K granularity
  long long* tl = (g_fwd_timeline && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 &&
                   (warp == 5 || warp == 0))
                      ? g_fwd_timeline
                      : nullptr;

  pdl_launch_dependents();
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < kFwdStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_consumed, 128);
    mbar_init(p_full, 128);
    mbar_init(p_free, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tP = tmem_base + 128, tO = tmem_base + 192;
  pdl_wait();  // qkv (the previous GEMM's output) is complete and visible

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTq * kHd * 2);
      tma_load_3d(sQ, &map_qkv, q_full, h * kHd, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kFwdStages;
        const uint32_t ph = ((j / kFwdStages) & 1) ^ 1;
        mbar_wait(&k_empty[st], ph);
        mbar_arrive_expect_tx(&k_full[st], kTk * kHd * 2);
        tma_load_3d(sK + st * (kTk * kHd * 2), &map_qkv, &k_full[st], p.D + h * kHd, j * kTk, b);
        mbar_wait(&v_empty[st], ph);
        mbar_arrive_expect_tx(&v_full[st], kTk * kHd * 2);
        tma_load_3d(sV + st * (kTk * kHd * 2), &map_qkv, &v_full[st], 2 * p.D + h * kHd, j * kTk, b);
      }
    }
  } else if (warp == 5) {
    // whole warp walks the loop; one elected lane issues MMAs / commits inside warp-uniform control flow (under
    // `if (lane == 0)` every UTCHMMA gets wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop)
    constexpr uint32_t idesc_s = make_idesc_bf16(kTq, kTk, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(kTq, kHd, 0, 1);  // B = V, MN-major
    const uint64_t dq = make_desc_kmajor(smem_u32(sQ));
    const uint64_t dk0 = make_desc_kmajor(smem_u32(sK));
    const uint64_t dv0 = make_desc_mnmajor(smem_u32(sV), 64 * 128);
    // ragged last KV tile: S only for the first nt16 keys (N of the MMA), PV only over those nt16 / 16 K steps
    const uint32_t idesc_s_tail = make_idesc_bf16(kTq, nt16, 0, 0);
    mbar_wait(q_full, 0);
    for (int j = -1; j < n_kv; ++j) {
      if (j + 1 < n_kv) {  // S_{j+1}: as soon as S_j sits in the softmax threads' registers
        const int st1 = (j + 1) % kFwdStages;
        if (j >= 0) mbar_wait(s_consumed, j & 1);
        mbar_wait(&k_full[st1], ((j + 1) / kFwdStages) & 1);
        tc_fence_after();
        const uint64_t dk = dk0 + static_cast<uint64_t>(st1 * (kTk * kHd * 2 >> 4));
        const uint32_t id = (j + 1 == n_kv - 1) ? idesc_s_tail : idesc_s;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kHd / 16; ++k) umma_ss(tS, dq + 2 * k, dk + 2 * k, id, k ? 1u : 0u);
          umma_commit(s_full);
          umma_commit(&k_empty[st1]);
        }
        __syncwarp();
      }
      TLF(0, j + 1, 0);
      if (j >= 0) {  // O (TMEM) += P_j V_j ; the softmax warps have rescaled O beforehand when the row max moved
        const int st = j % kFwdStages;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], (j / kFwdStages) & 1);
        tc_fence_after();
        const uint64_t dv = dv0 + static_cast<uint64_t>(st * (kTk * kHd * 2 >> 4));
        const int ksteps = (j == n_kv - 1) ? nt16 / 16 : kTk / 16;
        if (elect_one()) {
          umma_ts(tO, tP, dv, idesc_pv, j ? 1u : 0u);
#pragma unroll
          for (int k = 1; k < kTk / 16; ++k)
            if (k < ksteps) umma_ts(tO, tP + 8 * k, dv + 128 * k, idesc_pv, 1u);
          umma_commit(p_free);
          umma_commit(&v_empty[st]);
        }
        __syncwarp();
      }
      TLF(0, j + 1, 1);
    }
  } else {
    // ------------------------------ softmax warpgroup ------------------------------
    // Thread t owns query row t: one TMEM read of the 128 scores, running max with lazy rescaling
    // (O and l are only rescaled when the max grows by more than 2^8, so p <= 256 stays exact enough
    // in bf16/fp32 and the TMEM round trip for O is rare), P back to TMEM as the A operand of PV.
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int row = q0 + warp * 32 + lane;
    float m = -INFINITY, l = 0.f;

    const FwdBars fb{s_full, s_consumed, p_full, p_free};
    for (int j = 0; j < n_kv - 1; ++j)
      fwd_softmax_tile<4, PK, PM>(j, p.L, fb, tS, tP, tO, lane_base, p.sl2, m, l, tl);
    switch ((kv_valid_tail + 31) >> 5) {  // the last tile, by the number of 32-key chunks that hold keys
      case 1: fwd_softmax_tile<1, PK, PM>(n_kv - 1, p.L, fb, tS, tP, tO, lane_base, p.sl2, m, l, tl); break;
      case 2: fwd_softmax_tile<2, PK, PM>(n_kv - 1, p.L, fb, tS, tP, tO, lane_base, p.sl2, m, l, tl); break;
      case 3: fwd_softmax_tile<3, PK, PM>(n_kv - 1, p.L, fb, tS, tP, tO, lane_base, p.sl2, m, l, tl); break;
      default: fwd_softmax_tile<4, PK, PM>(n_kv - 1, p.L, fb, tS, tP, tO, lane_base, p.sl2, m, l, tl); break;
    }

    mbar_wait(p_free, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    // tcgen05.ld is warp-collective: every lane reads its O row, only valid rows are stored
    uint32_t o[kHd];
    tmem_ld32(tO + lane_base, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
    tmem_ld32(tO + lane_base + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
    tmem_ld_wait();
    if (row < p.L) {
      __nv_bfloat16* dst = p.o + (static_cast<size_t>(b) * p.L + row) * p.D + h * kHd;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int i = 0; i < kHd / 8; ++i) {
        uint4 v;
        v.x = pack_bf16(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
        v.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
        v.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
        v.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
        d4[i] = v;
      }
      p.lse2[(static_cast<size_t>(b) * p.H + h) * p.Lp + row] = m + log2f(l);
    } else {
      // pad rows [L, Lp): +inf makes the backward's exp2(s - lse2) vanish without a mask
      p.lse2[(static_cast<size_t>(b) * p.H + h) * p.Lp + row] = INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int attn_fwd(const void* qkv, void* o, float* lse2, int B, int L, int H, float scale, cudaStream_t st, int q_tiles) {
  if (B <= 0 || L <= 0 || H <= 0) return set_error(DCV_ERR_INVALID, "attn_fwd: empty problem");
  const int D = H * kHd;
  CUtensorMap map;
  if (int e = make_tmap_bf16_3d(&map, qkv, (uint64_t)3 * D, (uint64_t)L, (uint64_t)B, (uint64_t)3 * D * 2,
                                (uint64_t)L * 3 * D * 2, kHd, kTq, 1))
    return e;
  using KernelFn = void (*)(const CUtensorMap, const AttnFwdParams);
  KernelFn kern;
  switch (g_fwd_mode) {
    case 0: kern = attn_fwd_kernel<false, 0>; break;
    case 1: kern = attn_fwd_kernel<true, 0>; break;
    case 2: kern = attn_fwd_kernel<true, 0x88>; break;  // 2 of 8 pairs on the FMA pipe (default: fastest measured)
    case 3: kern = attn_fwd_kernel<true, 0xA4>; break;  // 3 of 8
    default: kern = attn_fwd_kernel<true, 0xAA>; break; // 4 of 8
  }
  DCV_TRY_SMEM_ATTR(kern, kFwdSmem);
  AttnFwdParams p;
  p.B = B; p.L = L; p.H = H; p.D = D;
  p.Lp = (L + 127) / 128 * 128;
  p.sl2 = scale * 1.4426950408889634f;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.lse2 = lse2;
  const int all_tiles = (L + kTq - 1) / kTq;
  dim3 grid(q_tiles > 0 && q_tiles < all_tiles ? q_tiles : all_tiles, H, B);
  ProfScope prof(PT_ATTN_FWD, st);
  DCV_CUDA(launch_pdl(kern, grid, dim3(192), kFwdSmem, st, map, p));
  count_launch();
  return 0;
}


}  // namespace dcv
