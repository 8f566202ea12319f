// Flash-attention backward on tcgen05 / TMEM (head_dim 64), replacing the autograd backward of
// reference models/vit.py:126-141 (softmax(q k^T * scale) v).
//
//   prep   : delta[b,h,q] = sum_d dO[q,d] * O[q,d] (0 on the pad rows), dq accumulator cleared
//   main   : one CTA per (128-key tile, head, image), 1 CTA / SM, looping over 128-query tiles i:
//              S_i^T  = K Q_i^T          dP_i^T = V dO_i^T                  (SS MMAs, operands K-major)
//              P_i^T  = exp2(S_i^T*sl2 - lse2[q])      -> bf16 into TMEM (A operand of the dV MMA)
//              dS_i^T = P_i^T o (dP_i^T - delta[q])    -> bf16 into smem (hand-swizzled UMMA tile)
//              dV  += P_i^T dO_i (TS MMA)   dK += dS_i^T Q_i (SS)   dQ_i = dS_i K   (SS, A MN-major)
//            Software pipeline: in phase i the compute threads produce P_i (MUFU-bound) and dS_{i-1} (FMA pipe)
//            in the same instruction stream; S_{i+1} / dP_i are issued while phase i is still running (as soon as
//            their TMEM inputs sit in registers), dK_{i-1} / dV_i / dQ_{i-1} right after it.
//            dQ_i is drained TMEM -> swizzled smem -> cp.reduce.async.bulk.tensor (fp32 add) into the [B,H,L,64]
//            accumulator; the softmax scale is folded into the dK epilogue and the finish kernel.
//   finish : scale * dq_acc fp32 [B,H,L,64] -> bf16 dqkv[:, :, 0:D]
//
// Warp roles (512 threads, registers re-balanced with setmaxnreg): warps 0-7 = two compute warpgroups, warpgroup g
// owns the query columns [64g, 64g+64) of every S^T / dP^T tile (TMEM lane = key row = 32*(warp%4)+lane);
// warp 8 = TMA producer; warp 9 = front MMA issuer (S, dP; owns TMEM); warp 10 = back MMA issuer (dK, dV, dQ);
// warps 12-15 drain dQ.
// TMEM columns: S^T [0,128) | dP^T [128,256) | dV [256,320) | dK [320,384) | dQ fp32 / P^T bf16 (time-shared) [384,448)
//               | K bf16 [448,480) | V bf16 [480,512)   (K, V: A operands of the S^T / dP^T TS-MMAs, copied once per CTA)
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
#ifdef DCV_ATTN_TIMELINE
#define TL(role, it, pt)                                                                      \
  do {                                                                                        \
    if (tl) tl[(role) * 1024 + (it) * 8 + (pt)] = clock64();                                   \
  } while (0)
#else
#define TL(role, it, pt) ((void)tl)
#endif

struct AttnBwdParams {
  int B, L, H, D, Lp;
  int n_q;      // query tiles to visit (all, or 1 when only the CLS rows carry gradient)
  float sl2;    // scale * log2(e)
  float scale;
  const float* lse2;   // [B,H,Lp]
  const float* delta;  // [B,H,Lp]
  __nv_bfloat16* dqkv; // [B,L,3D]
  float* dbias;        // optional [3D]: column sums of dqkv (qkv bias gradient) are ADDED here (dK / dV parts by the
                       // main kernel's epilogue, the dQ part by the finish kernel)
};

// smem: K,V | Q x3 stages | dO x2 stages | dS^T x2 buffers (each 2 chunks of [128 kv][64 q]) |
//       dQ staging (2 boxes of [128 q][32] fp32) | lse2 / delta of the query tile x3 stages (TMA bulk copies riding
//       on the Q barrier)
constexpr int kQStages = 3;
constexpr int kStatBytes = 2 * kTq * 4;  // 128 lse2 + 128 delta
constexpr int kBwdSmem = 2 * kTile16K + kQStages * kTile16K + 2 * kTile16K + 4 * kTile16K + 2 * kTile16K +
                         kQStages * kStatBytes + 1024 + 256;
// warpgroups 0-3: compute | warpgroup 4: TMA warp, two MMA-issuer warps, 1 idle | warpgroup 5: dQ drain
constexpr int kBwdThreads = 768;
constexpr int kWarpTma = 16, kWarpFront = 17, kWarpBack = 18, kWarpDrain0 = 20;

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

struct BwdBars {
  uint64_t kv_full, q_full[kQStages], q_empty[kQStages], do_full[2], do_empty[2];
  uint64_t s_full, dp_full, s_consumed, dp_consumed, phase_done, p_free, ds_free[2];
  uint64_t dq_full, dq_empty, dkv_full, kv_tmem;
  uint32_t tmem_slot;
};

// One software-pipelined phase of a compute thread (key row r, query columns [32cg, 32cg+32) of the tile):
//   A part (tile i)  : P^T = exp2(S^T * sl2 - lse2[q])                   -> bf16 into TMEM at the end of the phase
//   B part (tile i-1): dS^T = P^T o (dP^T - delta[q]), P^T kept packed in registers -> bf16 into the swizzled smem tile
// The B part's FMA-pipe work fills the issue slots the A part leaves while it waits on the MUFU; four compute warps
// per SM sub-partition hide the dependent-instruction latency.
// PK: packed-pair fp32 math (FFMA2 / FADD2 / FMUL2); PM: 8-bit mask over every group of 8 score pairs whose exponentials
// run on the FMA pipe (exp2_poly_f32x2) instead of the MUFU.
template <bool HAS_A, bool HAS_B, bool PK, int PM>
__device__ __forceinline__ void bwd_phase(int i, int cg, int r, uint32_t lane_base, BwdBars* bars, uint8_t* sStat,
                                          uint8_t* sdS, uint32_t tS, uint32_t tdP, uint32_t tP, float sl2, bool zero_row,
                                          uint32_t (&pkeep)[16], long long* tl) {
  uint32_t s[32];
  TL(1 + cg, i, 0);
  if (HAS_A) {
    mbar_wait(&bars->q_full[i % kQStages], (i / kQStages) & 1);  // lse2 / delta landed (long before S_i)
    mbar_wait(&bars->s_full, i & 1);
    tc_fence_after();
    tmem_ld32(tS + lane_base + cg * 32, s);
    tmem_ld_wait();
    tc_fence_before();
    mbar_arrive(&bars->s_consumed);  // S_{i+1} may overwrite tS
  }
  TL(1 + cg, i, 1);
  uint32_t sdS_g = 0, s_del = 0;
  if (HAS_B) {
    const int j = i - 1;
    mbar_wait(&bars->dp_full, j & 1);
    if (j >= 2) mbar_wait(&bars->ds_free[j & 1], ((j >> 1) - 1) & 1);  // dK_{j-2}, dQ_{j-2} have read this buffer
    tc_fence_after();
    sdS_g = smem_u32(sdS + (j & 1) * (2 * kTile16K) + (cg >> 1) * kTile16K);
    s_del = smem_u32(sStat + (j % kQStages) * kStatBytes) + kTq * 4 + cg * 128;
  }
  TL(1 + cg, i, 2);
  const uint32_t s_lse = smem_u32(sStat + (i % kQStages) * kStatBytes) + cg * 128;
#pragma unroll
  for (int c = 0; c < 2; ++c) {  // 16 query columns per step
    uint32_t d[16];
    if (HAS_B) tmem_ld16(tdP + lane_base + cg * 32 + 16 * c, d);
    if (HAS_A) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 l4 = ld_shared_f4(s_lse + (16 * c + 4 * q) * 4);
        if constexpr (PK) {
          const uint64_t sl2x2 = pack_f32x2(sl2, sl2);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int e = 16 * c + 4 * q + 2 * hh;
            const uint64_t x = fma_f32x2(pack_f32x2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), sl2x2,
                                         hh ? pack_f32x2(-l4.z, -l4.w) : pack_f32x2(-l4.x, -l4.y));
            float p0, p1;
            if ((PM >> ((e >> 1) & 7)) & 1) {
              unpack_f32x2(exp2_poly_f32x2(x), p0, p1);
            } else {
              unpack_f32x2(x, p0, p1);
              p0 = fast_exp2(p0);
              p1 = fast_exp2(p1);
            }
            s[e] = __float_as_uint(p0);
            s[e + 1] = __float_as_uint(p1);
          }
        } else {
          s[16 * c + 4 * q + 0] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(s[16 * c + 4 * q + 0]), sl2, -l4.x)));
          s[16 * c + 4 * q + 1] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(s[16 * c + 4 * q + 1]), sl2, -l4.y)));
          s[16 * c + 4 * q + 2] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(s[16 * c + 4 * q + 2]), sl2, -l4.z)));
          s[16 * c + 4 * q + 3] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(s[16 * c + 4 * q + 3]), sl2, -l4.w)));
        }
      }
    }
    if (HAS_B) {
      tmem_ld_wait();
      if (c == 1) {
        tc_fence_before();
        mbar_arrive(&bars->dp_consumed);  // dP_i may overwrite tdP
      }
#pragma unroll
      for (int v = 0; v < 2; ++v) {  // 8 queries -> one 16-byte slot of the swizzled dS^T tile
        const float4 da = ld_shared_f4(s_del + (16 * c + 8 * v) * 4), db = ld_shared_f4(s_del + (16 * c + 8 * v + 4) * 4);
        const float2 p0 = unpack_bf16(pkeep[8 * c + 4 * v + 0]), p1 = unpack_bf16(pkeep[8 * c + 4 * v + 1]);
        const float2 p2 = unpack_bf16(pkeep[8 * c + 4 * v + 2]), p3 = unpack_bf16(pkeep[8 * c + 4 * v + 3]);
        float e0, e1, e2, e3, e4, e5, e6, e7;
        if constexpr (PK) {
          unpack_f32x2(mul_f32x2(pack_f32x2(p0.x, p0.y),
                                 sub_f32x2(pack_f32x2(__uint_as_float(d[8 * v + 0]), __uint_as_float(d[8 * v + 1])),
                                           pack_f32x2(da.x, da.y))), e0, e1);
          unpack_f32x2(mul_f32x2(pack_f32x2(p1.x, p1.y),
                                 sub_f32x2(pack_f32x2(__uint_as_float(d[8 * v + 2]), __uint_as_float(d[8 * v + 3])),
                                           pack_f32x2(da.z, da.w))), e2, e3);
          unpack_f32x2(mul_f32x2(pack_f32x2(p2.x, p2.y),
                                 sub_f32x2(pack_f32x2(__uint_as_float(d[8 * v + 4]), __uint_as_float(d[8 * v + 5])),
                                           pack_f32x2(db.x, db.y))), e4, e5);
          unpack_f32x2(mul_f32x2(pack_f32x2(p3.x, p3.y),
                                 sub_f32x2(pack_f32x2(__uint_as_float(d[8 * v + 6]), __uint_as_float(d[8 * v + 7])),
                                           pack_f32x2(db.z, db.w))), e6, e7);
        } else {
          e0 = p0.x * (__uint_as_float(d[8 * v + 0]) - da.x);
          e1 = p0.y * (__uint_as_float(d[8 * v + 1]) - da.y);
          e2 = p1.x * (__uint_as_float(d[8 * v + 2]) - da.z);
          e3 = p1.y * (__uint_as_float(d[8 * v + 3]) - da.w);
          e4 = p2.x * (__uint_as_float(d[8 * v + 4]) - db.x);
          e5 = p2.y * (__uint_as_float(d[8 * v + 5]) - db.y);
          e6 = p3.x * (__uint_as_float(d[8 * v + 6]) - db.z);
          e7 = p3.y * (__uint_as_float(d[8 * v + 7]) - db.w);
        }
        st_shared_v4(sdS_g + sw128_offset(r, 4 * (cg & 1) + 2 * c + v), pack_bf16(e0, e1), pack_bf16(e2, e3),
                     pack_bf16(e4, e5), pack_bf16(e6, e7));
      }
    }
  }
  TL(1 + cg, i, 3);
  if (HAS_A) {
    // P^T shares its TMEM columns with the dQ accumulator: dV_{i-1} must have read P_{i-1} and the drain warps
    // must have taken dQ_{i-2} out before P_i goes in (dQ_{i-1} is issued behind dV_i by the same thread)
    if (i > 0) mbar_wait(&bars->p_free, (i - 1) & 1);
    if (i > 1) mbar_wait(&bars->dq_empty, (i - 2) & 1);
    if (i > 0) tc_fence_after();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      pkeep[j] = zero_row ? 0u : pack_bf16(__uint_as_float(s[2 * j]), __uint_as_float(s[2 * j + 1]));
    tmem_st16(tP + lane_base + cg * 16, pkeep);
    tmem_st_wait();
  }
  if (HAS_B) fence_proxy_async_smem();
  // last phase: the back issuer must have consumed phase_done(i-1) before this phase completes the barrier again
  // (mbarrier waits only tell the current phase from the previous one)
  if (!HAS_A) mbar_wait(&bars->p_free, (i - 1) & 1);
  tc_fence_before();
  mbar_arrive(&bars->phase_done);
  TL(1 + cg, i, 4);
}

template <bool PK, int PM>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                const __grid_constant__ CUtensorMap map_dq, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kTile16K;
  uint8_t* sQ = sV + kTile16K;               // 3 stages
  uint8_t* sdO = sQ + kQStages * kTile16K;   // 2 stages
  uint8_t* sdS = sdO + 2 * kTile16K;         // [2 buffers][2 q-chunks][128 kv rows][128 B] swizzled
  uint8_t* sdQ = sdS + 4 * kTile16K;         // [2 d-halves][128 q rows][32 fp32] swizzled
  uint8_t* sStat = sdQ + 2 * kTile16K;       // [3 stages][lse2 128 | delta 128] fp32
  BwdBars* bars = reinterpret_cast<BwdBars*>(sStat + kQStages * kStatBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * kTk;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_q = p.n_q;
  long long* tl = (g_attn_timeline && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 &&
                   (warp == kWarpFront || warp == kWarpBack || warp == 0 || warp == 4))
                      ? g_attn_timeline
                      : nullptr;

  pdl_launch_dependents();
  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_dq);
    mbar_init(&bars->kv_full, 1);
    for (int i = 0; i < kQStages; ++i) {
      mbar_init(&bars->q_full[i], 1);
      mbar_init(&bars->q_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->do_full[i], 1);
      mbar_init(&bars->do_empty[i], 2);  // dP_i (front issuer) and dV_i (back issuer) have both read dO_i
      mbar_init(&bars->ds_free[i], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->dp_full, 1);
    mbar_init(&bars->s_consumed, 512);
    mbar_init(&bars->dp_consumed, 512);
    mbar_init(&bars->phase_done, 512);
    mbar_init(&bars->p_free, 1);
    mbar_init(&bars->dq_full, 1);
    mbar_init(&bars->dq_empty, 128);  // the four drain warps
    mbar_init(&bars->dkv_full, 1);
    mbar_init(&bars->kv_tmem, 512);
    fence_barrier_init();
  }
  if (warp == kWarpFront) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  // P^T (bf16) and the dQ accumulator time-share columns [384,448); K and V sit in TMEM as A operands of S^T / dP^T
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384, tP = tmem_base + 384, tK = tmem_base + 448, tV = tmem_base + 480;
  pdl_wait();  // dO / delta (projection dgrad) and the cleared dQ accumulator are complete and visible

  if (warp >= kWarpTma && warp < kWarpDrain0) {
    setmaxnreg_dec<40>();
    if (warp == kWarpTma) {
      // ------------------------------------ TMA producer ------------------------------------
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars->kv_full, 2 * kTile16K);
        tma_load_3d(sK, &map_qkv, &bars->kv_full, p.D + h * kHd, kv0, b);
        tma_load_3d(sV, &map_qkv, &bars->kv_full, 2 * p.D + h * kHd, kv0, b);
        for (int i = 0; i < n_q; ++i) {
          const int sq = i % kQStages, sd = i & 1;
          if (i >= kQStages) mbar_wait(&bars->q_empty[sq], ((i / kQStages) - 1) & 1);
          mbar_arrive_expect_tx(&bars->q_full[sq], kTile16K + kStatBytes);
          tma_load_3d(sQ + sq * kTile16K, &map_qkv, &bars->q_full[sq], h * kHd, i * kTq, b);
          const size_t so = (static_cast<size_t>(b) * p.H + h) * p.Lp + static_cast<size_t>(i) * kTq;
          bulk_load_1d(sStat + sq * kStatBytes, p.lse2 + so, kTq * 4, &bars->q_full[sq]);
          bulk_load_1d(sStat + sq * kStatBytes + kTq * 4, p.delta + so, kTq * 4, &bars->q_full[sq]);
          if (i >= 2) mbar_wait(&bars->do_empty[sd], ((i >> 1) - 1) & 1);
          mbar_arrive_expect_tx(&bars->do_full[sd], kTile16K);
          tma_load_3d(sdO + sd * kTile16K, &map_do, &bars->do_full[sd], h * kHd, i * kTq, b);
        }
      }
    } else if (warp == kWarpFront) {
      // ---------------------------------- front MMA issuer ----------------------------------
      // S_{i+1} as soon as S_i sits in the compute threads' registers, dP_i as soon as dP_{i-1} has been read.
      // The whole warp walks the loop (all lanes wait on the barriers) and ONE elected lane issues the MMAs /
      // commits inside warp-uniform control flow: under `if (lane == 0)` the compiler has to assume divergent
      // operands and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop.
      // A single warp runs its dependent instruction stream at ~5 clk per instruction, so the issue work of one
      // (q-tile, k-tile) pair is split over two warps (this one and the back issuer, warp 10).
      constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);   // S^T, dP^T
      const uint64_t dQ_k0 = make_desc_kmajor(smem_u32(sQ));     // + stage * (16 KB >> 4)
      const uint64_t dO_k0 = make_desc_kmajor(smem_u32(sdO));
      mbar_wait(&bars->kv_tmem, 0);  // K, V copied into TMEM by the compute warps
      tc_fence_after();
      for (int i = -1; i < n_q; ++i) {
        if (i + 1 < n_q) {
          const int s1 = (i + 1) % kQStages;
          if (i >= 0) mbar_wait(&bars->s_consumed, i & 1);
          mbar_wait(&bars->q_full[s1], ((i + 1) / kQStages) & 1);
          tc_fence_after();
          const uint64_t dQn_k = dQ_k0 + static_cast<uint64_t>(s1 * (kTile16K >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ts(tS, tK + 8 * k, dQn_k + 2 * k, id_s, k ? 1u : 0u);
            umma_commit(&bars->s_full);
          }
          __syncwarp();
        }
        TL(0, i + 1, 0);
        if (i >= 0) {
          if (i > 0) mbar_wait(&bars->dp_consumed, (i - 1) & 1);
          mbar_wait(&bars->do_full[i & 1], (i >> 1) & 1);
          tc_fence_after();
          const uint64_t dO_k = dO_k0 + static_cast<uint64_t>((i & 1) * (kTile16K >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ts(tdP, tV + 8 * k, dO_k + 2 * k, id_s, k ? 1u : 0u);
            umma_commit(&bars->dp_full);
            umma_commit(&bars->do_empty[i & 1]);
          }
          __syncwarp();
        }
        TL(0, i + 1, 1);
      }
    } else if (warp == kWarpBack) {
      // ----------------------------------- back MMA issuer -----------------------------------
      // After phase i (P_i in TMEM, dS_{i-1} in smem): dK_{i-1} (frees the Q stage), dV_i (frees P and the dO
      // stage), dQ_{i-1}.
      constexpr uint32_t id_kv = make_idesc_bf16(128, 64, 0, 1);   // dV, dK : A K-major, B MN-major
      constexpr uint32_t id_dq = make_idesc_bf16(128, 64, 1, 1);   // dQ     : A MN-major, B MN-major
      const uint64_t dK_mn = make_desc_mnmajor(smem_u32(sK), kTile16K);
      const uint64_t dQ_mn0 = make_desc_mnmajor(smem_u32(sQ), kTile16K);
      const uint64_t dO_mn0 = make_desc_mnmajor(smem_u32(sdO), kTile16K);
      const uint64_t dS_k00 = make_desc_kmajor(smem_u32(sdS));
      const uint64_t dS_mn0 = make_desc_mnmajor(smem_u32(sdS), kTile16K);
      mbar_wait(&bars->kv_full, 0);
      for (int i = 0; i <= n_q; ++i) {
        const int j = i - 1;  // tile whose dS^T was produced in phase i
        mbar_wait(&bars->phase_done, i & 1);
        tc_fence_after();
        TL(3, i, 0);
        if (i > 0) {  // dK += dS_j^T Q_j
          const uint64_t dS_k0 = dS_k00 + static_cast<uint64_t>((j & 1) * (2 * kTile16K >> 4));
          const uint64_t dS_k1 = dS_k0 + (kTile16K >> 4);
          const uint64_t dQ_mn = dQ_mn0 + static_cast<uint64_t>((j % kQStages) * (kTile16K >> 4));
          if (elect_one()) {
            umma_ss(tdK, dS_k0, dQ_mn, id_kv, j ? 1u : 0u);
#pragma unroll
            for (int k = 1; k < 8; ++k)
              umma_ss(tdK, (k < 4 ? dS_k0 : dS_k1) + 2 * (k & 3), dQ_mn + 128 * k, id_kv, 1u);
            umma_commit(&bars->q_empty[j % kQStages]);
          }
          __syncwarp();
        }
        if (i < n_q) {  // dV += P_i^T dO_i  (A = P^T from TMEM: 16 q per K step = 8 packed columns)
          const uint64_t dO_mn = dO_mn0 + static_cast<uint64_t>((i & 1) * (kTile16K >> 4));
          if (elect_one()) {
            umma_ts(tdV, tP, dO_mn, id_kv, i ? 1u : 0u);
#pragma unroll
            for (int k = 1; k < 8; ++k) umma_ts(tdV, tP + 8 * k, dO_mn + 128 * k, id_kv, 1u);
            umma_commit(&bars->p_free);
            umma_commit(&bars->do_empty[i & 1]);
          }
          __syncwarp();
        }
        TL(3, i, 1);
        if (i > 0) {  // dQ_j = dS_j K
          if (j > 0) {  // the drain warps have read dQ_{j-1} out of TMEM
            mbar_wait(&bars->dq_empty, (j - 1) & 1);
            tc_fence_after();
          }
          const uint64_t dS_mn = dS_mn0 + static_cast<uint64_t>((j & 1) * (2 * kTile16K >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_ss(tdQ, dS_mn + 128 * k, dK_mn + 128 * k, id_dq, k ? 1u : 0u);
            umma_commit(&bars->dq_full);
            umma_commit(&bars->ds_free[j & 1]);
          }
          __syncwarp();
        }
        TL(3, i, 2);
      }
      if (elect_one()) umma_commit(&bars->dkv_full);
      __syncwarp();
    }
  } else if (warp >= kWarpDrain0) {
    setmaxnreg_dec<56>();
    // ------------------------------------ dQ drain warps ------------------------------------
    // dQ_i (128 queries x 64) : TMEM -> two swizzled [128][32] fp32 boxes in smem -> TMA reduce-add into the
    // [B,H,L,64] accumulator.  Off the critical path of the compute warpgroups.
    const int q4 = warp & 3;  // TMEM lane quadrant
    const int r = q4 * 32 + lane;
    const int tid_d = threadIdx.x - kWarpDrain0 * 32;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    const uint32_t sdQ0 = smem_u32(sdQ), sdQ1 = smem_u32(sdQ + kTile16K);
    for (int i = 0; i < n_q; ++i) {
      mbar_wait(&bars->dq_full, i & 1);
      tc_fence_after();
      // two 32-column halves, one after the other (keeps this warpgroup at 56 registers).  Vectorised
      // red.global.add.v4.f32 straight from registers would skip the staging traffic in shared memory (this kernel is
      // bound by shared-memory bandwidth) but was measured 25 % slower end to end (16-byte L2 reductions).
      uint32_t qr[32];
      tmem_ld32(tdQ + lane_base, qr);
      tmem_ld_wait();
      if (tid_d == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous reduce has read the boxes
      asm volatile("bar.sync 3, 128;" ::: "memory");
#pragma unroll
      for (int v = 0; v < 8; ++v) st_shared_v4(sdQ0 + sw128_offset(r, v), qr[4 * v], qr[4 * v + 1], qr[4 * v + 2], qr[4 * v + 3]);
      tmem_ld32(tdQ + lane_base + 32, qr);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bars->dq_empty);
#pragma unroll
      for (int v = 0; v < 8; ++v) st_shared_v4(sdQ1 + sw128_offset(r, v), qr[4 * v], qr[4 * v + 1], qr[4 * v + 2], qr[4 * v + 3]);
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
    setmaxnreg_inc<96>();
    // --------------------------------- compute warpgroups ---------------------------------
    const int cg = warp >> 2;                 // 32-column group of the query tile
    const int q4 = warp & 3;                  // TMEM lane quadrant
    const int r = q4 * 32 + lane;             // key row
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    const bool kv_ok = kv0 + r < p.L;

    {  // K (column groups 0, 1) and V (2, 3) -> TMEM: 16 packed columns = 32 head-dim elements of row r per thread
      mbar_wait(&bars->kv_full, 0);
      const uint32_t src = smem_u32((cg < 2) ? sK : sV);
      uint32_t w[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 x = ld_shared_f4(src + sw128_offset(r, 4 * (cg & 1) + c));
        w[4 * c] = __float_as_uint(x.x); w[4 * c + 1] = __float_as_uint(x.y);
        w[4 * c + 2] = __float_as_uint(x.z); w[4 * c + 3] = __float_as_uint(x.w);
      }
      tmem_st16(((cg < 2) ? tK : tV) + lane_base + 16 * (cg & 1), w);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars->kv_tmem);
    }
    uint32_t pkeep[16];  // P^T of the previous tile, packed bf16 (B part of the next phase)
    bwd_phase<true, false, PK, PM>(0, cg, r, lane_base, bars, sStat, sdS, tS, tdP, tP, p.sl2, !kv_ok, pkeep, tl);
    for (int i = 1; i < n_q; ++i)
      bwd_phase<true, true, PK, PM>(i, cg, r, lane_base, bars, sStat, sdS, tS, tdP, tP, p.sl2, !kv_ok, pkeep, tl);
    bwd_phase<false, true, PK, PM>(n_q, cg, r, lane_base, bars, sStat, sdS, tS, tdP, tP, p.sl2, !kv_ok, pkeep, tl);

    // ---- epilogue: dK (x scale) and dV rows of this key tile; column group cg writes d columns [16cg, 16cg+16) ----
    mbar_wait(&bars->dkv_full, 0);
    tc_fence_after();
#pragma unroll
    for (int which = 0; which < 2; ++which) {  // 0: dK -> column block D, 1: dV -> column block 2D
      const uint32_t tsrc = which == 0 ? tdK : tdV;
      const float mul = which == 0 ? p.scale : 1.0f;
      uint32_t a[16];
      tmem_ld16(tsrc + lane_base + cg * 16, a);
      tmem_ld_wait();
      if (p.dbias != nullptr) {
        // column sums over the 32 key rows of this warp (rows beyond L hold exact zeros): transposing butterfly,
        // 16 shuffles, after which lane l owns column 8*b4 + 4*b3 + 2*b2 + b1 of its bits; even lanes add it
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(a[i]) * mul;
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1) {
          const bool up = (lane & (2 * w)) != 0;
#pragma unroll
          for (int i = 0; i < w; ++i) {
            const float send = up ? x[i] : x[i + w];
            const float keep = up ? x[i + w] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * w);
          }
        }
        x[0] += __shfl_xor_sync(0xffffffffu, x[0], 1);
        if ((lane & 1) == 0) {
          const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
          atomicAdd(p.dbias + (which + 1) * p.D + h * kHd + cg * 16 + col, x[0]);
        }
      }
      if (kv_ok) {
        __nv_bfloat16* dst = p.dqkv + (static_cast<size_t>(b) * p.L + kv0 + r) * (3 * p.D) + (which + 1) * p.D +
                             h * kHd + cg * 16;
        uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
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
  if (warp == kWarpFront) {
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

// scale * dq_acc fp32 [B,H,L,64] -> dqkv bf16 [B,L,3D] columns [h*64, h*64+64); optionally the column sums of the
// result (the q third of the qkv bias gradient) are added to dbias[0, D).
// One CTA per (128 query rows, b*H + h): thread = (row lane 0..31, 8-column group); 4 rows per thread.
constexpr int kFinRows = 128;
__global__ void __launch_bounds__(256)
attn_bwd_finish_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias,
                       int B, int L, int H, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float part[8][kHd];
  const int bh = blockIdx.x;
  const int bb = bh / H, hh = bh - bb * H;
  const int c8 = threadIdx.x & 7, rr = threadIdx.x >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * kHd;
  float sacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int it = 0; it < kFinRows / 32; ++it) {
    const int q = blockIdx.y * kFinRows + it * 32 + rr;
    if (q < L) {
      const float4* src = reinterpret_cast<const float4*>(dq_acc + (static_cast<size_t>(bh) * L + q) * kHd + c8 * 8);
      const float4 a = __ldg(src), c = __ldg(src + 1);
      const float v[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, c.x * scale, c.y * scale, c.z * scale, c.w * scale};
      uint4 o;
      o.x = pack_bf16(v[0], v[1]);
      o.y = pack_bf16(v[2], v[3]);
      o.z = pack_bf16(v[4], v[5]);
      o.w = pack_bf16(v[6], v[7]);
      *reinterpret_cast<uint4*>(dqkv + (static_cast<size_t>(bb) * L + q) * (3 * D) + hh * kHd + c8 * 8) = o;
#pragma unroll
      for (int i = 0; i < 8; ++i) sacc[i] += v[i];
    }
  }
  if (dbias == nullptr) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // the warp's 4 row lanes
    sacc[i] += __shfl_xor_sync(0xffffffffu, sacc[i], 8);
    sacc[i] += __shfl_xor_sync(0xffffffffu, sacc[i], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) part[warp][c8 * 8 + i] = sacc[i];
  }
  __syncthreads();
  if (threadIdx.x < kHd) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
    atomicAdd(dbias + hh * kHd + threadIdx.x, t);
  }
}

}  // namespace

static int g_bwd_mode = 1;
void debug_set_attn_bwd_mode(int m) { g_bwd_mode = m; }
int debug_fwd_timeline(long long* buf);  // attention.cu
int debug_attn_timeline(long long* buf) {
  DCV_CUDA(cudaMemcpyToSymbol(g_attn_timeline, &buf, sizeof(buf)));
  return debug_fwd_timeline(buf ? buf + 4096 : nullptr);
}

int attn_bwd(const void* qkv, const void* o, const void* dO, const float* lse2, float* delta, float* dq_acc,
             void* dqkv, int B, int L, int H, float scale, cudaStream_t st, bool cls_only, bool delta_ready,
             float* dbias_qkv, bool dq_cleared) {
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
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnBwdParams);
  KernelFn kern;
  switch (g_bwd_mode) {
    case 0: kern = attn_bwd_kernel<false, 0>; break;
    case 1: kern = attn_bwd_kernel<true, 0>; break;
    case 2: kern = attn_bwd_kernel<true, 0x80>; break;  // 1 of 8 pairs on the FMA pipe
    default: kern = attn_bwd_kernel<true, 0x88>; break; // 2 of 8
  }
  DCV_TRY_SMEM_ATTR(kern, kBwdSmem);
  {
    ProfScope prof(PT_ATTN_BWD_PREP, st);
    if (delta_ready) {
      // produced by the projection-dgrad GEMM epilogue
    } else if (cls_only) {
      attn_bwd_prep_cls_kernel<<<B * H, 32, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(o),
                                                     reinterpret_cast<const __nv_bfloat16*>(dO), delta, B, L, H, Lp);
    } else {
      const long long total = static_cast<long long>(B) * L * H * 8;
      attn_bwd_prep_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
          reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(dO), delta, B, L, H, Lp);
    }
    DCV_CUDA(cudaGetLastError());
    if (!dq_cleared) DCV_CUDA(cudaMemsetAsync(dq_acc, 0, static_cast<size_t>(B) * H * L * kHd * sizeof(float), st));
  }
  AttnBwdParams p;
  p.B = B; p.L = L; p.H = H; p.D = D; p.Lp = Lp;
  p.n_q = cls_only ? 1 : (L + kTq - 1) / kTq;
  p.scale = scale;
  p.sl2 = scale * 1.4426950408889634f;
  p.lse2 = lse2; p.delta = delta;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.dbias = dbias_qkv;
  dim3 grid((L + kTk - 1) / kTk, H, B);
  {
    ProfScope prof(PT_ATTN_BWD, st);
    DCV_CUDA(launch_pdl(kern, grid, dim3(kBwdThreads), kBwdSmem, st, map_qkv, map_do, map_dq, p));
  }
  {
    ProfScope prof(PT_ATTN_BWD_FIN, st);
    DCV_CUDA(launch_pdl(attn_bwd_finish_kernel, dim3(B * H, (L + kFinRows - 1) / kFinRows), dim3(256), 0, st, dq_acc,
                        reinterpret_cast<__nv_bfloat16*>(dqkv), dbias_qkv, B, L, H, scale));
  }
  count_launch(delta_ready ? 2 : 3);
  return 0;
}

}  // namespace dcv
