// Patch embedding of the channel-adaptive ViT as ONE kernel, fed by TMA straight from the fp32 image
// (reference models/dichavit.py:210 channel gather, :377 Conv3d(1, D, (1, P, P)) per channel, :378-389 the per-token
// normalisation of TDL, :409-411 + :561-565 bias / channel token / positional embedding):
//
//   TMA   5-D box (x = W, r = 4 pixel rows, py = strips, c = idx[c'], b) of the fp32 image -> shared memory
//   conv  4 warps: fp32 pixels -> [hi | lo] bf16 split, written as SWIZZLE_128B K-major UMMA operand tiles (and the hi
//         part to global memory: the A operand of the backward's weight-gradient GEMM)
//   MMA   tcgen05: tokens[128, 384] += A_hi Whi^T + A_lo Whi^T + A_hi Wlo^T   (= x W^T to ~2^-16, fp32 in TMEM;
//         the split weight [Whi | Whi | Wlo] of embed.cu streams through a 2-stage TMA ring, 192 output columns a time)
//   epi   4 warps, one token row per thread: tokens = acc + (bias + channel token + pos) -> global fp32;
//         |y|^2 of y = acc + bias, rnorm = 1 / max(|y|, eps); second TMEM pass: f = y * rnorm summed over the tile's
//         tokens (32 x 32 transposing butterfly per warp) -> fp32 atomics into S[b, c', :], Q[b, c'] -- what
//         tdl_sum_kernel produced from a second pass over the tokens in global memory.
//
// Replaces im2col_gather_kernel + the EPI_EMBED GEMM + tdl_sum_kernel (about 51 MB image -> 77 MB patches -> 77 MB
// re-read -> 77 MB tokens -> 77 MB re-read at the JUMP-CP shape) by 51 MB in, 77 MB tokens + 26 MB hi-patches out.
// One CTA per (image, sampled channel, group of `spt` patch rows): 7 x 14 = 98 tokens of a 224 x 224 plane.
// Shapes: P = 16, D = 384 (the accumulator row has to fit 512 TMEM columns for the row norm), fp32 input; everything
// else keeps the three-kernel path (embed_fused_ok()).
#include <algorithm>

#include "common.cuh"
#include "host.h"

namespace dcv {

namespace {

constexpr int kP = 16;
constexpr int kK = kP * kP;          // 256
constexpr int kD = 384;
constexpr int kChunks = kK / 64;     // 4 K chunks of 64 = 4 pixel rows of the patch
constexpr int kATile = 128 * 64 * 2; // 16 KB
constexpr int kBHalf = 192;
constexpr int kBTile = kBHalf * 64 * 2;  // 24 KB
constexpr int kStageMax = 25600;     // fp32 staging buffer of one K chunk (spt strips x 4 rows x W)
constexpr int kEfThreads = 384;
constexpr int kEfSmem = 2 * 2 * kATile + 2 * 2 * kBTile + 2 * kStageMax + 1024 + 256;
static_assert(kEfSmem <= 227 * 1024, "shared memory budget");

struct EfBars {
  uint64_t stage_full[2], stage_empty[2], a_full[2], a_empty[2], b_full[2], b_empty[2], acc_full;
  uint32_t tmem_slot;
};

struct EfParams {
  int B, C, Cs, H, W;
  int wp, hp;           // patches per row / column of the image
  int spt, tpp;         // patch rows (strips) per tile, tiles per (image, channel) plane
  int N, T;             // tokens per plane, per image
  const int* idx;       // [Cs] source channel of every sampled channel (NULL = identity)
  const float* bias;    // [D]
  const float* addend;  // [T, D]  bias + channel token + positional embedding of every token
  float* tokens;        // [B, T + 1, D]
  __nv_bfloat16* patches;  // [B * T, 3K], only the first K columns (hi part) are written
  float* S;             // [B, Cs, D]   (zeroed by the host)
  float* Q;             // [B, Cs]
  float* rnorm;         // [B, T]
  int tdl_on;
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}

// debug timeline: when non-null, CTA `kTlCta` records clock64() stamps (tools/embed_timeline.py)
__device__ long long* g_ef_timeline = nullptr;
constexpr int kTlCta = 200;
#define EF_TL(slot) do { if (tl) tl[slot] = clock64(); } while (0)

__global__ void __launch_bounds__(kEfThreads, 1)
embed_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const EfParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                         // [2 buffers][hi | lo] x 16 KB
  uint8_t* sB = sA + 4 * kATile;              // [2 stages][Whi | Wlo] x 24 KB
  uint8_t* sStage = sB + 4 * kBTile;          // [2 buffers] x kStageMax
  EfBars* bars = reinterpret_cast<EfBars*>(sStage + 2 * kStageMax);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = blockIdx.x;
  const int g = tile % p.tpp;
  tile /= p.tpp;
  const int cs = tile % p.Cs;
  const int b = tile / p.Cs;
  const int strips = min(p.spt, p.hp - g * p.spt);   // patch rows of this tile that exist
  const int ntok = strips * p.wp;                    // <= 128
  const int t0 = cs * p.N + g * p.spt * p.wp;        // first token of the tile inside its image
  const uint32_t stage_bytes = static_cast<uint32_t>(p.spt) * 4u * p.W * 4u;
  long long* tl = (g_ef_timeline && blockIdx.x == kTlCta && lane == 0) ? g_ef_timeline : nullptr;
  if (warp == 0) EF_TL(70);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->stage_full[i], 1);
      mbar_init(&bars->stage_empty[i], 128);
      mbar_init(&bars->a_full[i], 128);
      mbar_init(&bars->a_empty[i], 1);
      mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->b_empty[i], 1);
    }
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  if (warp == 0) EF_TL(71);
  pdl_wait();
  if (warp == 0) EF_TL(72);

  if (warp == 0) {
    // ---- image staging: one 5-D box per K chunk (4 pixel rows of every patch row of the tile) ----
    if (lane == 0) {
      const int c_src = p.idx ? __ldg(p.idx + cs) : cs;
      for (int kc = 0; kc < kChunks; ++kc) {
        const int sb = kc & 1;
        if (kc >= 2) mbar_wait(&bars->stage_empty[sb], 0);
        EF_TL(kc);
        mbar_arrive_expect_tx(&bars->stage_full[sb], stage_bytes);
        tma_load_5d(sStage + sb * kStageMax, &map_x, &bars->stage_full[sb], 0, 4 * kc, g * p.spt, c_src, b);
      }
    }
  } else if (warp == 3) {
    // ---- split weight: [Whi | Wlo] columns of K chunk kc, output columns [192 h, 192 h + 192) ----
    if (lane == 0) {
      for (int it = 0; it < 2 * kChunks; ++it) {
        const int kc = it >> 1, h = it & 1;
        if (kc >= 1) mbar_wait(&bars->b_empty[h], (kc - 1) & 1);
        EF_TL(8 + it);
        mbar_arrive_expect_tx(&bars->b_full[h], 2 * kBTile);
        tma_load_2d(sB + h * (2 * kBTile), &map_w, &bars->b_full[h], kc * 64, h * kBHalf);
        tma_load_2d(sB + h * (2 * kBTile) + kBTile, &map_w, &bars->b_full[h], 2 * kK + kc * 64, h * kBHalf);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    constexpr uint32_t idesc = make_idesc_bf16(128, kBHalf, 0, 0);
    for (int it = 0; it < 2 * kChunks; ++it) {
      const int kc = it >> 1, h = it & 1, sb = kc & 1;
      if (h == 0) mbar_wait(&bars->a_full[sb], (kc >> 1) & 1);
      if (h == 0) EF_TL(88 + kc);
      mbar_wait(&bars->b_full[h], kc & 1);
      tc_fence_after();
      EF_TL(16 + it);
      const uint64_t da_hi = make_desc_kmajor(smem_u32(sA + sb * (2 * kATile)));
      const uint64_t da_lo = make_desc_kmajor(smem_u32(sA + sb * (2 * kATile) + kATile));
      const uint64_t db_hi = make_desc_kmajor(smem_u32(sB + h * (2 * kBTile)));
      const uint64_t db_lo = make_desc_kmajor(smem_u32(sB + h * (2 * kBTile) + kBTile));
      const uint32_t d_tmem = tmem_base + h * kBHalf;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_ss(d_tmem, da_hi + 2 * ks, db_hi + 2 * ks, idesc, (kc | ks) ? 1u : 0u);
          umma_ss(d_tmem, da_lo + 2 * ks, db_hi + 2 * ks, idesc, 1u);
          umma_ss(d_tmem, da_hi + 2 * ks, db_lo + 2 * ks, idesc, 1u);
        }
        umma_commit(&bars->b_empty[h]);
        if (h == 1) umma_commit(&bars->a_empty[sb]);
        if (it == 2 * kChunks - 1) umma_commit(&bars->acc_full);
      }
      __syncwarp();
      EF_TL(24 + it);
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- conversion: staged fp32 pixels -> bf16 hi / lo operand tiles (+ hi part to global memory) ----
    const int tid_c = threadIdx.x - 128;
    const size_t row_g0 = static_cast<size_t>(b) * p.T + t0;
    for (int kc = 0; kc < kChunks; ++kc) {
      const int sb = kc & 1;
      mbar_wait(&bars->stage_full[sb], (kc >> 1) & 1);
      if (warp == 4) EF_TL(32 + kc * 3);
      if (kc >= 2) mbar_wait(&bars->a_empty[sb], 0);
      if (warp == 4) EF_TL(33 + kc * 3);
      const uint32_t s_src = smem_u32(sStage + sb * kStageMax);
      const uint32_t s_hi = smem_u32(sA + sb * (2 * kATile)), s_lo = s_hi + kATile;
      for (int task = tid_c; task < ntok * 8; task += 128) {
        const int t = task >> 3, k8 = task & 7;
        const int pyl = t / p.wp, px = t - pyl * p.wp;
        const uint32_t src = s_src + (((pyl * 4 + (k8 >> 1)) * p.W + px * kP + (k8 & 1) * 8) << 2);
        const float4 v0 = ld_shared_f4(src), v1 = ld_shared_f4(src + 16);
        uint4 hi, lo;
        hi.x = pack_bf16(v0.x, v0.y); hi.y = pack_bf16(v0.z, v0.w); hi.z = pack_bf16(v1.x, v1.y); hi.w = pack_bf16(v1.z, v1.w);
        const float2 h0 = unpack_bf16(hi.x), h1 = unpack_bf16(hi.y), h2 = unpack_bf16(hi.z), h3 = unpack_bf16(hi.w);
        lo.x = pack_bf16(v0.x - h0.x, v0.y - h0.y); lo.y = pack_bf16(v0.z - h1.x, v0.w - h1.y);
        lo.z = pack_bf16(v1.x - h2.x, v1.y - h2.y); lo.w = pack_bf16(v1.z - h3.x, v1.w - h3.y);
        const uint32_t off = sw128_offset(t, k8);
        st_shared_v4(s_hi + off, hi.x, hi.y, hi.z, hi.w);
        st_shared_v4(s_lo + off, lo.x, lo.y, lo.z, lo.w);
        *reinterpret_cast<uint4*>(p.patches + (row_g0 + t) * (3 * kK) + kc * 64 + k8 * 8) = hi;
      }
      fence_proxy_async_smem();
      if (warp == 4) EF_TL(34 + kc * 3);
      mbar_arrive(&bars->a_full[sb]);
      mbar_arrive(&bars->stage_empty[sb]);
    }
  }
  if (warp >= 4) {
    // ---- epilogue: warps 4-11 (the conversion warps join once their last chunk is converted) ----
    // Two warps per TMEM lane quadrant, taking alternate 32-column chunks.  A thread owns one token ROW in TMEM, the
    // global tensors want one ROW per warp instruction: every chunk goes through a private swizzled [32][32] fp32 tile
    // in shared memory (the idle staging buffers) -- written row-wise (4 wavefronts per STS.128), read column-wise
    // (1 wavefront per LDS.32) -- so that the addend loads and token stores are full 128-byte lines.  (The first
    // version read / wrote global memory row-per-thread: 32 lines per instruction, and the epilogue was 75 % of the
    // kernel.)
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const int m = q * 32 + lane;
    const bool valid = m < ntok;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* epi = sStage;                                       // reused: the main loop is over
    const uint32_t s_tile = smem_u32(epi) + (warp - 4) * 4096;   // private [32 rows][128 B] tile
    float* s_bias = reinterpret_cast<float*>(epi + 32768);       // [384]
    float* s_ss = reinterpret_cast<float*>(epi + 32768 + 1536);  // [2 groups][128 rows]
    mbar_wait(&bars->acc_full, 0);                               // all MMAs done: accumulators final, smem idle
    tc_fence_after();
    const int tb = warp == 4 ? 48 : 96;  // timeline rows of one conversion + epilogue warp and one epilogue-only warp
    if (warp == 4 || warp == 8) EF_TL(tb);
    for (int i = threadIdx.x - 128; i < kD; i += 256) s_bias[i] = __ldg(p.bias + i);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (warp == 4 || warp == 8) EF_TL(tb + 1);
    const int rows_w = max(0, min(32, ntok - q * 32));           // token rows of this warp that exist
    const size_t tok0 = static_cast<size_t>(b) * (p.T + 1) + 1 + t0 + q * 32;  // global token row of the warp's row 0
    const float* add0 = p.addend + static_cast<size_t>(t0 + q * 32) * kD;
    float ss = 0.f;
    for (int c = grp; c < kD / 32; c += 2) {
      // the chunk's addend values (lane = column, one per token row) are requested first: 32 independent coalesced
      // loads whose L2 latency hides behind the TMEM read and the row-wise part below
      float ad[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) ad[r] = r < rows_w ? __ldg(add0 + static_cast<size_t>(r) * kD + 32 * c + lane) : 0.f;
      uint32_t v[32];
      tmem_ld32(taddr + 32 * c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bi = *reinterpret_cast<const float4*>(s_bias + 32 * c + 4 * j);
        const float y0 = __uint_as_float(v[4 * j]) + bi.x, y1 = __uint_as_float(v[4 * j + 1]) + bi.y,
                    y2 = __uint_as_float(v[4 * j + 2]) + bi.z, y3 = __uint_as_float(v[4 * j + 3]) + bi.w;
        if (valid) ss += (y0 * y0 + y1 * y1) + (y2 * y2 + y3 * y3);
        st_shared_v4(s_tile + sw128_offset(lane, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      __syncwarp();
      // transposed: lane = column; one token row (128 contiguous bytes of addend / tokens) per iteration
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        float a;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(s_tile + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2)));
        if (r < rows_w) p.tokens[(tok0 + r) * kD + 32 * c + lane] = a + ad[r];
      }
      __syncwarp();
      if (warp == 4 || warp == 8) EF_TL(tb + 2 + (c >> 1));
    }
    if (p.tdl_on) {
      s_ss[grp * 128 + m] = ss;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 4 || warp == 8) EF_TL(tb + 8);
      const float sst = s_ss[m] + s_ss[128 + m];
      const float inv = valid ? 1.0f / fmaxf(sqrtf(sst), 1e-12f) : 0.f;
      if (valid && grp == 0) p.rnorm[static_cast<size_t>(b) * p.T + t0 + m] = inv;
      float* s_row = p.S + (static_cast<size_t>(b) * p.Cs + cs) * kD;
      for (int c = grp; c < kD / 32; c += 2) {
        uint32_t v[32];
        tmem_ld32(taddr + 32 * c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bi = *reinterpret_cast<const float4*>(s_bias + 32 * c + 4 * j);
          // (rows beyond ntok hold whatever the unused operand rows produced: select, do not multiply by 0)
          const float f0 = valid ? (__uint_as_float(v[4 * j]) + bi.x) * inv : 0.f;
          const float f1 = valid ? (__uint_as_float(v[4 * j + 1]) + bi.y) * inv : 0.f;
          const float f2 = valid ? (__uint_as_float(v[4 * j + 2]) + bi.z) * inv : 0.f;
          const float f3 = valid ? (__uint_as_float(v[4 * j + 3]) + bi.w) * inv : 0.f;
          st_shared_v4(s_tile + sw128_offset(lane, j), __float_as_uint(f0), __float_as_uint(f1), __float_as_uint(f2),
                       __float_as_uint(f3));
        }
        __syncwarp();
        float colsum = 0.f;  // rows beyond ntok were written with inv = 0
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          float a;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(s_tile + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2)));
          colsum += a;
        }
        atomicAdd(s_row + 32 * c + lane, colsum);
        __syncwarp();
        if (warp == 4 || warp == 8) EF_TL(tb + 9 + (c >> 1));
      }
      if (grp == 0) {
        const float qp = warp_sum(valid ? sst * inv * inv : 0.f);
        if (lane == 0) atomicAdd(p.Q + static_cast<size_t>(b) * p.Cs + cs, qp);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) EF_TL(73);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// v2: the same arithmetic as a PERSISTENT kernel (one CTA per SM walking tiles blockIdx.x, + gridDim.x, ...) whose
// stages run ahead across tile boundaries.  The clock64 timeline of the one-tile-per-CTA kernel above
// (tools/embed_timeline.py, profiles/r2_embed_fused_timeline.txt) shows a 40 k-cycle tile of which the epilogue is
// 24 k with nothing else in flight, and every CTA of a wave in the same phase at the same time (all SMs read, then
// all SMs write).  Here
//   * the image / weight TMA producers and the conversion warps work on tile i+1 while the 8 epilogue warps drain
//     tile i (the only serialisation left is the accumulator: the 384 fp32 columns of a tile cannot be
//     double-buffered in 512 TMEM columns, so the MMAs of tile i+1 wait for the last tcgen05.ld of tile i);
//   * the weight ring is 3 stages of [Whi | Wlo] x 96 output columns (24 KB) that stream continuously -- the split
//     weight does not depend on the tile;
//   * the conversion has a fixed per-thread task list (no integer division in the loop, 8 independent tasks in
//     flight, quarter-warps read 256 contiguous bytes of the staged image row instead of four rows 896 B apart);
//   * the TDL column sums are reduced in registers (32 x 32 transposing butterfly, 31 SHFL) instead of a second trip
//     through shared memory, so the second TMEM pass is bound by tcgen05.ld.
constexpr int kV2Threads = 512;
constexpr int kBN = 96;                      // output columns of one weight stage / one MMA
constexpr int kBTile2 = kBN * 64 * 2;        // 12 KB
constexpr int kBStage2 = 2 * kBTile2;        // [Whi | Wlo]
constexpr int kBStages = 3;
constexpr int kNQ = kD / kBN;                // 4 column groups
constexpr int kEpiTile = 4096;               // private [32][32] fp32 transposition tile of an epilogue warp
constexpr int kV2Misc = 4096;                // bias, row-norm exchange, barriers
constexpr int kV2Smem = 4 * kATile + kBStages * kBStage2 + 2 * kStageMax + 8 * kEpiTile + kV2Misc + 1024;
static_assert(kV2Smem <= 227 * 1024, "shared memory budget");
static_assert(kD % kBN == 0 && kBN % 16 == 0, "weight stage shape");

struct Ef2Bars {
  uint64_t stage_full[2], stage_empty[2], a_full[2], a_empty[2], b_full[kBStages], b_empty[kBStages], acc_full, acc_empty;
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(kV2Threads, 1)
embed_fused_v2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const EfParams p,
                      const int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // [2 buffers][hi | lo] x 16 KB
  uint8_t* sB = sA + 4 * kATile;                        // [3 stages][Whi | Wlo] x 12 KB
  uint8_t* sStage = sB + kBStages * kBStage2;           // [2 buffers] x kStageMax
  uint8_t* sEpi = sStage + 2 * kStageMax;               // [8 warps] x 4 KB
  float* s_bias = reinterpret_cast<float*>(sEpi + 8 * kEpiTile);  // [384]
  float* s_ss = s_bias + kD;                            // [2 groups][128 rows]
  Ef2Bars* bars = reinterpret_cast<Ef2Bars*>(s_ss + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage_bytes = static_cast<uint32_t>(p.spt) * 4u * p.W * 4u;
  // debug timeline of CTA 5: slot = role * 128 + 16 * (tile iteration) + point, roles 0 = MMA, 1 = conversion warp 4,
  // 2 = epilogue warp 8 (tools/embed_timeline.py)
  long long* tl = (g_ef_timeline && blockIdx.x == 5 && lane == 0) ? g_ef_timeline : nullptr;
  if (tl && warp == 0) tl[511] = clock64();

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->stage_full[i], 1);
      mbar_init(&bars->stage_empty[i], 128);
      mbar_init(&bars->a_full[i], 128);
      mbar_init(&bars->a_empty[i], 1);
    }
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->b_empty[i], 1);
    }
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->acc_empty, 256);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_slot, 512);
  pdl_wait();
  for (int i = threadIdx.x; i < kD; i += kV2Threads) s_bias[i] = __ldg(p.bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == 0) {
    // ---- image staging: one 5-D box per K chunk (4 pixel rows of every patch row of the tile), two boxes in flight ----
    if (lane == 0) {
      uint32_t n = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int g = tile % p.tpp, bc = tile / p.tpp, cs = bc % p.Cs, b = bc / p.Cs;
        const int c_src = p.idx ? __ldg(p.idx + cs) : cs;
        for (int kc = 0; kc < kChunks; ++kc, ++n) {
          const uint32_t sb = n & 1, use = n >> 1;
          if (use >= 1) mbar_wait(&bars->stage_empty[sb], (use - 1) & 1);
          mbar_arrive_expect_tx(&bars->stage_full[sb], stage_bytes);
          tma_load_5d(sStage + sb * kStageMax, &map_x, &bars->stage_full[sb], 0, 4 * kc, g * p.spt, c_src, b);
        }
      }
    }
  } else if (warp == 3) {
    // ---- split weight: [Whi | Wlo] columns of K chunk kc, output columns [96 q, 96 q + 96); the same 16 stages per tile ----
    if (lane == 0) {
      uint32_t m = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int it = 0; it < kChunks * kNQ; ++it, ++m) {
          const int kc = it / kNQ, q = it % kNQ;
          const uint32_t s = m % kBStages, ub = m / kBStages;
          if (ub >= 1) mbar_wait(&bars->b_empty[s], (ub - 1) & 1);
          mbar_arrive_expect_tx(&bars->b_full[s], kBStage2);
          tma_load_2d(sB + s * kBStage2, &map_w, &bars->b_full[s], kc * 64, q * kBN);
          tma_load_2d(sB + s * kBStage2 + kBTile2, &map_w, &bars->b_full[s], 2 * kK + kc * 64, q * kBN);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    constexpr uint32_t idesc = make_idesc_bf16(128, kBN, 0, 0);
    uint32_t n = 0, m = 0, j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      if (tl && j < 8) tl[16 * j] = clock64();
      if (j >= 1) mbar_wait(&bars->acc_empty, (j - 1) & 1);   // the epilogue has read the previous tile out of TMEM
      tc_fence_after();
      if (tl && j < 8) tl[16 * j + 1] = clock64();
      for (int kc = 0; kc < kChunks; ++kc, ++n) {
        const uint32_t sb = n & 1, use = n >> 1;
        mbar_wait(&bars->a_full[sb], use & 1);
        if (tl && j < 8) tl[16 * j + 2 + 2 * kc] = clock64();
        const uint64_t da_hi = make_desc_kmajor(smem_u32(sA + sb * (2 * kATile)));
        const uint64_t da_lo = make_desc_kmajor(smem_u32(sA + sb * (2 * kATile) + kATile));
        for (int q = 0; q < kNQ; ++q, ++m) {
          const uint32_t s = m % kBStages, ub = m / kBStages;
          mbar_wait(&bars->b_full[s], ub & 1);
          tc_fence_after();
          const uint64_t db_hi = make_desc_kmajor(smem_u32(sB + s * kBStage2));
          const uint64_t db_lo = make_desc_kmajor(smem_u32(sB + s * kBStage2 + kBTile2));
          const uint32_t d_tmem = tmem_base + q * kBN;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_ss(d_tmem, da_hi + 2 * ks, db_hi + 2 * ks, idesc, (kc | ks) ? 1u : 0u);
              umma_ss(d_tmem, da_lo + 2 * ks, db_hi + 2 * ks, idesc, 1u);
              umma_ss(d_tmem, da_hi + 2 * ks, db_lo + 2 * ks, idesc, 1u);
            }
            umma_commit(&bars->b_empty[s]);
            if (q == kNQ - 1) umma_commit(&bars->a_empty[sb]);
            if (q == kNQ - 1 && kc == kChunks - 1) umma_commit(&bars->acc_full);
          }
          __syncwarp();
        }
        if (tl && j < 8) tl[16 * j + 3 + 2 * kc] = clock64();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- conversion: staged fp32 pixels -> bf16 hi / lo operand tiles (+ hi part to global memory) ----
    // lane = (half of a 16-pixel patch row, 4 consecutive tokens, 4 pixel rows): a quarter-warp reads 256 contiguous
    // bytes of one staged image row.  Thread's tokens: t = 16 i + 4 wc + tq, i = 0..7 (geometry is tile-independent).
    const int wc = warp - 4;
    const int half = lane & 1, tq = (lane >> 1) & 3, rr = lane >> 3;
    const int k8 = rr * 2 + half;
    uint32_t soff[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = 16 * i + 4 * wc + tq;
      const int pyl = t / p.wp, px = t - pyl * p.wp;
      soff[i] = static_cast<uint32_t>(((pyl * 4 + rr) * p.W + px * kP + half * 8) << 2);
    }
    uint32_t n = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int g = tile % p.tpp, bc = tile / p.tpp, cs = bc % p.Cs, b = bc / p.Cs;
      const int strips = min(p.spt, p.hp - g * p.spt);
      const int ntok = strips * p.wp;
      const size_t row_g0 = static_cast<size_t>(b) * p.T + cs * p.N + g * p.spt * p.wp;
      for (int kc = 0; kc < kChunks; ++kc, ++n) {
        const uint32_t sb = n & 1, use = n >> 1;
        mbar_wait(&bars->stage_full[sb], use & 1);
        if (tl && warp == 4 && n < 32) tl[128 + 4 * n] = clock64();
        if (use >= 1) mbar_wait(&bars->a_empty[sb], (use - 1) & 1);
        if (tl && warp == 4 && n < 32) tl[128 + 4 * n + 1] = clock64();
        const uint32_t s_src = smem_u32(sStage + sb * kStageMax);
        const uint32_t s_hi = smem_u32(sA + sb * (2 * kATile)), s_lo = s_hi + kATile;
        __nv_bfloat16* gdst = p.patches + row_g0 * (3 * kK) + kc * 64 + k8 * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int t = 16 * i + 4 * wc + tq;
          if (t < ntok) {
            const float4 v0 = ld_shared_f4(s_src + soff[i]), v1 = ld_shared_f4(s_src + soff[i] + 16);
            uint4 hi, lo;
            hi.x = pack_bf16(v0.x, v0.y); hi.y = pack_bf16(v0.z, v0.w); hi.z = pack_bf16(v1.x, v1.y); hi.w = pack_bf16(v1.z, v1.w);
            const float2 h0 = unpack_bf16(hi.x), h1 = unpack_bf16(hi.y), h2 = unpack_bf16(hi.z), h3 = unpack_bf16(hi.w);
            lo.x = pack_bf16(v0.x - h0.x, v0.y - h0.y); lo.y = pack_bf16(v0.z - h1.x, v0.w - h1.y);
            lo.z = pack_bf16(v1.x - h2.x, v1.y - h2.y); lo.w = pack_bf16(v1.z - h3.x, v1.w - h3.y);
            const uint32_t off = sw128_offset(t, k8);
            st_shared_v4(s_hi + off, hi.x, hi.y, hi.z, hi.w);
            st_shared_v4(s_lo + off, lo.x, lo.y, lo.z, lo.w);
            *reinterpret_cast<uint4*>(gdst + static_cast<size_t>(t) * (3 * kK)) = hi;
          }
        }
        fence_proxy_async_smem();
        if (tl && warp == 4 && n < 32) tl[128 + 4 * n + 2] = clock64();
        mbar_arrive(&bars->a_full[sb]);
        mbar_arrive(&bars->stage_empty[sb]);
      }
    }
  } else if (warp >= 8) {
    // ---- epilogue: warps 8-15, two per TMEM lane quadrant, alternate 32-column chunks (see the kernel above for the
    // transposition through a private swizzled tile that makes the addend loads / token stores full lines) ----
    const int we = warp - 8, q = we & 3, grp = we >> 2;
    const int m = q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t s_tile = smem_u32(sEpi) + we * kEpiTile;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const int g = tile % p.tpp, bc = tile / p.tpp, cs = bc % p.Cs, b = bc / p.Cs;
      const int strips = min(p.spt, p.hp - g * p.spt);
      const int ntok = strips * p.wp;
      const int t0 = cs * p.N + g * p.spt * p.wp;
      const bool valid = m < ntok;
      const int rows_w = max(0, min(32, ntok - q * 32));
      const size_t tok0 = static_cast<size_t>(b) * (p.T + 1) + 1 + t0 + q * 32;
      const float* add0 = p.addend + static_cast<size_t>(t0 + q * 32) * kD;
      // Transposed phase: lane = (row group rq, 16-byte column chunk cq); iteration i handles tile row 4 i + rq, so that one
      // LDS.128 / LDG.128 / STG.128 covers four full 128-byte token rows.  The addend values of a chunk are requested a
      // whole chunk ahead (the register that just delivered row 4 i + rq of chunk c is refilled for chunk c + 2, this
      // group's next one; the first chunk is requested before the accumulator is even complete): in the first version each
      // chunk exposed one full L2 round trip (2 750 cycles per chunk, 60 % of the kernel).
      const int rq = lane >> 3, cq = lane & 7;
      const float* add_l = add0 + static_cast<size_t>(rq) * kD + 4 * cq;
      float* tok_l = p.tokens + (tok0 + rq) * kD + 4 * cq;
      float4 ad[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        ad[i] = 4 * i + rq < rows_w ? __ldg(reinterpret_cast<const float4*>(add_l + static_cast<size_t>(4 * i) * kD + 32 * grp))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      if (tl && warp == 8 && j < 8) tl[256 + 16 * j] = clock64();
      mbar_wait(&bars->acc_full, j & 1);
      tc_fence_after();
      if (tl && warp == 8 && j < 8) tl[256 + 16 * j + 1] = clock64();
      float ss = 0.f;
      for (int c = grp; c < kD / 32; c += 2) {
        uint32_t v[32];
        tmem_ld32(taddr + 32 * c, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float4 bi = *reinterpret_cast<const float4*>(s_bias + 32 * c + 4 * jj);
          const float y0 = __uint_as_float(v[4 * jj]) + bi.x, y1 = __uint_as_float(v[4 * jj + 1]) + bi.y,
                      y2 = __uint_as_float(v[4 * jj + 2]) + bi.z, y3 = __uint_as_float(v[4 * jj + 3]) + bi.w;
          if (valid) ss += (y0 * y0 + y1 * y1) + (y2 * y2 + y3 * y3);
          st_shared_v4(s_tile + sw128_offset(lane, jj), v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = 4 * i + rq;
          const float4 a = ld_shared_f4(s_tile + row * 128 + ((cq ^ (row & 7)) << 4));
          if (row < rows_w) {
            *reinterpret_cast<float4*>(tok_l + static_cast<size_t>(4 * i) * kD + 32 * c) =
                make_float4(a.x + ad[i].x, a.y + ad[i].y, a.z + ad[i].z, a.w + ad[i].w);
            if (c + 2 < kD / 32)
              ad[i] = __ldg(reinterpret_cast<const float4*>(add_l + static_cast<size_t>(4 * i) * kD + 32 * (c + 2)));
          }
        }
        __syncwarp();
      }
      if (tl && warp == 8 && j < 8) tl[256 + 16 * j + 2] = clock64();
      if (p.tdl_on) {
        s_ss[grp * 128 + m] = ss;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tl && warp == 8 && j < 8) tl[256 + 16 * j + 3] = clock64();
        const float sst = s_ss[m] + s_ss[128 + m];
        const float inv = valid ? 1.0f / fmaxf(sqrtf(sst), 1e-12f) : 0.f;
        if (valid && grp == 0) p.rnorm[static_cast<size_t>(b) * p.T + t0 + m] = inv;
        float* s_row = p.S + (static_cast<size_t>(b) * p.Cs + cs) * kD;
        for (int c = grp; c < kD / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(taddr + 32 * c, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float4 bi = *reinterpret_cast<const float4*>(s_bias + 32 * c + 4 * jj);
            // (rows beyond ntok hold whatever the unused operand rows produced: select, do not multiply by 0)
            f[4 * jj] = valid ? (__uint_as_float(v[4 * jj]) + bi.x) * inv : 0.f;
            f[4 * jj + 1] = valid ? (__uint_as_float(v[4 * jj + 1]) + bi.y) * inv : 0.f;
            f[4 * jj + 2] = valid ? (__uint_as_float(v[4 * jj + 2]) + bi.z) * inv : 0.f;
            f[4 * jj + 3] = valid ? (__uint_as_float(v[4 * jj + 3]) + bi.w) * inv : 0.f;
          }
          // transposing butterfly: afterwards f[0] of lane l = sum over the warp's 32 token rows of column 32 c + l
#pragma unroll
          for (int s = 16; s >= 1; s >>= 1) {
            const bool up = (lane & s) != 0;
#pragma unroll
            for (int i = 0; i < s; ++i) {
              const float keep = up ? f[i + s] : f[i];
              const float send = up ? f[i] : f[i + s];
              f[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
            }
          }
          atomicAdd(s_row + 32 * c + lane, f[0]);
        }
        if (grp == 0) {
          const float qp = warp_sum(valid ? sst * inv * inv : 0.f);
          if (lane == 0) atomicAdd(p.Q + static_cast<size_t>(b) * p.Cs + cs, qp);
        }
      }
      if (tl && warp == 8 && j < 8) tl[256 + 16 * j + 4] = clock64();
      tc_fence_before();
      mbar_arrive(&bars->acc_empty);   // this thread's tcgen05.ld of the tile have completed
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int debug_embed_timeline(long long* buf) {
  DCV_CUDA(cudaMemcpyToSymbol(g_ef_timeline, &buf, sizeof(buf)));
  return 0;
}

constexpr int kEmbedFusedDefault = 2;
// 0 = three-kernel path, 1 = one tile per CTA, 2 = persistent pipelined kernel
static int g_embed_fused = -1;
void debug_set_embed_fused(int on) { g_embed_fused = on < 0 ? -1 : (on > 2 ? 2 : on); }  // < 0: back to the default

bool embed_fused_ok(const dcv_embed_dims& d, int x_is_u8) {
  if (g_embed_fused < 0) {
    const char* e = getenv("DCV_EMBED_FUSED");
    g_embed_fused = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : kEmbedFusedDefault;
  }
  if (!g_embed_fused || x_is_u8) return false;
  if (d.P != kP || d.D != kD || d.W % kP || d.H % kP || d.W > 256) return false;
  const int wp = d.W / kP, hp = d.H / kP;
  if (wp > 128) return false;
  const int tpp = (hp * wp + 127) / 128;
  const int spt = (hp + tpp - 1) / tpp;
  if (spt * wp > 128 || spt > 256) return false;
  if (static_cast<long long>(spt) * 4 * d.W * 4 > kStageMax) return false;
  return true;
}

int embed_fused_fwd(const dcv_embed_dims& d, const void* x, const int* idx, const void* wsplit, const float* bias,
                    const float* addend, float* tokens, void* patches, float* S, float* Q, float* rnorm, int tdl_on,
                    cudaStream_t st) {
  const int wp = d.W / kP, hp = d.H / kP, N = wp * hp, T = d.Cs * N;
  EfParams p;
  p.B = d.B; p.C = d.C; p.Cs = d.Cs; p.H = d.H; p.W = d.W;
  p.wp = wp; p.hp = hp;
  p.tpp = (N + 127) / 128;
  p.spt = (hp + p.tpp - 1) / p.tpp;
  p.N = N; p.T = T;
  p.idx = idx; p.bias = bias; p.addend = addend; p.tokens = tokens;
  p.patches = reinterpret_cast<__nv_bfloat16*>(patches);
  p.S = S; p.Q = Q; p.rnorm = rnorm; p.tdl_on = tdl_on;
  if (tdl_on) {
    if (!S || !Q || !rnorm) return set_error(DCV_ERR_INVALID, "embed_fused_fwd: TDL buffers missing");
    DCV_CUDA(cudaMemsetAsync(S, 0, static_cast<size_t>(d.B) * d.Cs * kD * sizeof(float), st));
    DCV_CUDA(cudaMemsetAsync(Q, 0, static_cast<size_t>(d.B) * d.Cs * sizeof(float), st));
  }
  CUtensorMap mx, mw;
  if (int e = make_tmap_f32_5d(&mx, x, (uint64_t)d.W, kP, (uint64_t)hp, (uint64_t)d.C, (uint64_t)d.B, (uint64_t)d.W * 4,
                               (uint64_t)kP * d.W * 4, (uint64_t)d.H * d.W * 4, (uint64_t)d.C * d.H * d.W * 4, d.W, 4,
                               p.spt, 1, 1))
    return e;
  const int n_tiles = d.B * d.Cs * p.tpp;
  if (g_embed_fused == 2) {
    if (int e = make_tmap_bf16_2d(&mw, wsplit, (uint64_t)3 * kK, (uint64_t)kD, (uint64_t)3 * kK * 2, 64, kBN)) return e;
    DCV_TRY_SMEM_ATTR(embed_fused_v2_kernel, kV2Smem);
    ProfScope prof(PT_EMBED_GEMM, st);
    DCV_CUDA(launch_pdl(embed_fused_v2_kernel, dim3(std::min(n_tiles, num_sms())), dim3(kV2Threads), kV2Smem, st, mx, mw, p,
                        n_tiles));
    count_launch();
    return 0;
  }
  if (int e = make_tmap_bf16_2d(&mw, wsplit, (uint64_t)3 * kK, (uint64_t)kD, (uint64_t)3 * kK * 2, 64, kBHalf)) return e;
  DCV_TRY_SMEM_ATTR(embed_fused_kernel, kEfSmem);
  ProfScope prof(PT_EMBED_GEMM, st);
  DCV_CUDA(launch_pdl(embed_fused_kernel, dim3(n_tiles), dim3(kEfThreads), kEfSmem, st, mx, mw, p));
  count_launch();
  return 0;
}

}  // namespace dcv
