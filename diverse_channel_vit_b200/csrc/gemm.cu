// tcgen05 GEMM family for the transformer blocks of the DiChaViT hot path.
//
//   gemm_nt :  C[M,N] = A[M,K] * B[N,K]^T  (+ fused epilogue)        -- both operands K-major
//              forward Linear layers (reference models/vit.py:116,118,71-73) and their
//              dgrad (dX = dY * W: the same kernel with W[out,in] consumed directly as an MN-major B operand,
//              template flag Bmn -- no transposed copy of W exists).
//   gemm_tn :  C[N,K] += A[M,N]^T * B[M,K]                             -- both operands MN-major
//              weight gradients dW = dY^T X, reduction over the B*L token rows, split-K.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B boxes, kStages-deep ring)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma, fp32 accumulators in TMEM)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue       (tcgen05.ld -> registers -> bias/GELU/residual -> swizzled smem -> TMA store)
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the
// main loop of tile i+1.
#include <cstdlib>
#include "common.cuh"
#include "host.h"

namespace dcv {

constexpr int kBM = 128;
constexpr int kBK = 64;  // one 128-byte swizzle atom of bf16

enum GemmEpilogue : int {
  EPI_BIAS = 0,        // out_bf16 = acc (+ bias)
  EPI_BIAS_GELU = 1,   // out_bf16 = h = acc + bias ; out2_bf16 = gelu(h)
  EPI_BIAS_RESID = 2,  // out_f32 = resid + acc + bias   (resid may alias out_f32)
  EPI_DGELU = 3,       // out_bf16 = acc * gelu'(aux)   (+ optional column sums of the result, see GemmNtParams::delta)
  EPI_F32 = 4,         // out_f32 = acc (+ bias)
  EPI_EMBED = 5,       // out_f32[(m / T) * L + 1 + m % T][:] = acc + addend[m % T][:]   (patch embedding)
  EPI_DELTA = 6,       // out_bf16 = dO = acc ; delta[b, h, q] = sum over head h's 64 columns of bf16(dO) * aux (aux = the
                       // attention output O): the softmax-backward row term, one column chunk == one head
};

struct GemmNtParams {
  int M, N, K;
  int ldo;   // leading dimension (elements) of out / out2
  int ldin;  // leading dimension of the epilogue input (resid / aux); == ldo unless rows are gathered
  const float* bias;
  __nv_bfloat16* out_bf16;
  __nv_bfloat16* out2_bf16;
  float* out_f32;
  const float* resid;
  const __nv_bfloat16* aux;
  // EPI_EMBED: rows m = (image b, token t) with t in [0,T); the output is the [B, L=T+1, D] token
  // tensor (row 0 of every image is the CLS token, written elsewhere); addend fp32 [T, ldo]
  int map_T, map_L;
  const float* addend;
  // EPI_DELTA: rows m = (image b, query q) with q in [0, seq_L); delta fp32 [B, N/64, seq_Lp]
  // EPI_DGELU: optional fp32 [N]: the column sums of the output (bias gradient of the layer below) are ADDED to it
  float* delta;
  int seq_L, seq_Lp;
};

// Epilogue staging: the 128 x BN accumulator tile leaves through shared memory in column chunks of
// 128 bytes per row (64 bf16 or 32 fp32 columns): each epilogue warp writes its 32 rows of a chunk into
// a private SWIZZLE_128B [32][128 B] buffer and issues its own TMA store (full-line, asynchronous,
// M-tail clipped by the tensor map).  Inputs of the epilogue (residual / GELU pre-activation) arrive
// the same way through per-warp TMA loads, fetched while the previous chunk is being computed.
constexpr int kChunkBytes = kBM * 128;  // 16 KB

template <int EPI>
struct EpiTraits {
  static constexpr bool kTma = EPI != EPI_EMBED;
  static constexpr bool kF32Out = EPI == EPI_BIAS_RESID || EPI == EPI_F32 || EPI == EPI_EMBED;
  static constexpr bool kHasIn = EPI == EPI_BIAS_RESID || EPI == EPI_DGELU || EPI == EPI_DELTA;
  static constexpr bool kTwoOut = EPI == EPI_BIAS_GELU;
  static constexpr int kCW = kF32Out ? 32 : 64;  // columns per chunk
  // out[2] (+ out2[2]) (+ in[2])
  static constexpr int kStagingBytes = kTma ? (2 + (kTwoOut ? 2 : 0) + (kHasIn ? 2 : 0)) * kChunkBytes : 0;
};

template <int BN, int EPI>
struct NtCfg {
  static constexpr int kStageBytes = (kBM + BN) * kBK * 2;
  static constexpr int kStagingBytes = EpiTraits<EPI>::kStagingBytes;
  static constexpr int kMaxStages = (232448 - 1024 - 256 - kStagingBytes) / kStageBytes;
  static constexpr int kStages = kMaxStages > 6 ? 6 : kMaxStages;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(kStages >= 3, "pipeline too shallow");
};

constexpr int kEpiWarps = 8;  // two warps per TMEM lane quadrant, taking alternate column chunks
constexpr int kNtThreads = 128 + kEpiWarps * 32;

// CM > 1: thread-block cluster of CM CTAs working on CM consecutive M-tiles of the same N-tile.  Every CTA loads
// its own A tile and 1/CM of the shared B tile, multicasting that slice into the shared memory of all CTAs of the
// cluster: L2 -> SM operand traffic per CTA drops from (128 + BN) to (128 + BN / CM) rows per K step (the L2
// bandwidth, ~6.3 KB/clk chip-wide, is what caps a 128 x 192 single-CTA tile at ~1 PFLOP/s).
// MODE 2 (pair MMA, `cta_group::2`): the two CTAs of a cluster compute ONE 256 x BN tile with a single MMA stream
// issued by the leader CTA; each CTA loads its 128 rows of A and HALF of B into its own shared memory and receives its
// 128 accumulator rows in its own TMEM.  Per-SM operand inflow per K step drops from (128 + BN) to (128 + BN/2) rows --
// the measured ~36-45 B/clk/SM L2 -> SM rate is what caps the single-CTA main loop.
template <int BN, int EPI, bool kBMn, int MODE>
__global__ void __launch_bounds__(kNtThreads, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_out2,
               const __grid_constant__ CUtensorMap map_in, const GemmNtParams p) {
  constexpr int CM = MODE == 0 ? 1 : 2;     // CTAs per cluster (consecutive M-tiles of one N-tile)
  constexpr bool kPair = MODE == 2;          // cta_group::2 MMA
  static_assert(!(kPair && kBMn), "pair MMA implemented for K-major B only");
  using Cfg = NtCfg<BN, EPI>;
  using ET = EpiTraits<EPI>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                              // kStages x [128][64] bf16
  uint8_t* smem_b = smem + kStages * (kBM * kBK * 2);  // kStages x [BN][64] bf16
  uint8_t* stage_out = smem + kStages * Cfg::kStageBytes;           // [2] x 16 KB
  uint8_t* stage_out2 = stage_out + 2 * kChunkBytes;                // [2] (GELU second output)
  uint8_t* stage_in = stage_out + (ET::kTwoOut ? 4 : 2) * kChunkBytes;  // [2] (residual / pre-activation)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint64_t* in_full = bars + 2 * kStages + 4;  // [kEpiWarps]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4 + kEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int n_tiles = p.N / BN;
  const int num_kb = (p.K + kBK - 1) / kBK;
  // work distribution: "cluster tiles" of CM consecutive M-tiles x one N-tile, M-major so that CTAs running at the
  // same time share A panels in L2; every CTA of a cluster runs the same number of iterations
  const int crank = CM > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CM;
  const int num_clusters = gridDim.x / CM;
  const int num_ct = ((m_tiles + CM - 1) / CM) * n_tiles;
  const int my_iters = cluster_id < num_ct ? (num_ct - cluster_id + num_clusters - 1) / num_clusters : 0;
  auto tile_coords = [&](int it, int& m0, int& n0) {
    const int ct = cluster_id + it * num_clusters;
    m0 = ((ct / n_tiles) * CM + crank) * kBM;  // may lie beyond M (odd tail): TMA zero-fills loads, clips stores
    n0 = (ct % n_tiles) * BN;
  };
  constexpr uint16_t kMcMask = static_cast<uint16_t>((1u << CM) - 1);
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (ET::kTma) tma_prefetch_desc(&map_out);
    if (ET::kTwoOut) tma_prefetch_desc(&map_out2);
    if (ET::kHasIn) tma_prefetch_desc(&map_in);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      // multicast mode: released by the MMA warp of every CTA that received the stage; pair mode: by the leader's
      // multicast commit only
      mbar_init(&empty_bar[i], kPair ? 1 : CM);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kPair ? 2 * kEpiWarps : kEpiWarps);  // one arrive per epilogue warp (of both CTAs)
    }
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(&in_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (kPair) tmem_alloc_pair(tmem_base_slot, Cfg::kTmemCols);
    else tmem_alloc(tmem_base_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  if (CM > 1) cluster_sync_all();  // peers' barriers must be initialised before remote arrives / multicast writes
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_wait();  // the previous kernel's outputs (our operands, residual, gradient accumulators) are complete and visible

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_iters; ++it) {
        int m0, n0;
        tile_coords(it, m0, n0);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sb = smem_b + stage * (BN * kBK * 2);
          if (kPair) {
            // both CTAs credit the LEADER's barrier: A tile + half of B from each of them
            constexpr uint32_t kHalfBytes = (kBM + BN / 2) * kBK * 2;
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kHalfBytes);
            tma_load_2d_pair(smem_a + stage * (kBM * kBK * 2), &map_a, &full_bar[stage], kb * kBK, m0);
            tma_load_2d_pair(sb, &map_b, &full_bar[stage], kb * kBK, n0 + crank * (BN / 2));
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          // bytes landing in THIS CTA's stage: own A tile + all CM slices of B
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * (kBM * kBK * 2), &map_a, &full_bar[stage], kb * kBK, m0);
          if (kBMn) {
            // B stored [K rows][N cols] (e.g. a weight W[out,in] used as dY*W): MN-major boxes of
            // [64 reduction rows][64 output columns]; a cluster slice is 64/CM reduction rows of every box
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) {
              if (CM > 1)
                tma_load_2d_mc(sb + c * (64 * 128) + crank * (kBK / CM) * 128, &map_b, &full_bar[stage], n0 + c * 64,
                               kb * kBK + crank * (kBK / CM), kMcMask);
              else
                tma_load_2d(sb + c * (64 * 128), &map_b, &full_bar[stage], n0 + c * 64, kb * kBK);
            }
          } else if (CM > 1) {
            tma_load_2d_mc(sb + crank * (BN / CM) * 128, &map_b, &full_bar[stage], kb * kBK, n0 + crank * (BN / CM),
                           kMcMask);
          } else {
            tma_load_2d(sb, &map_b, &full_bar[stage], kb * kBK, n0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    // whole warp walks the loop, one elected lane issues inside warp-uniform control flow (avoids the per-UTCHMMA
    // ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall the compiler emits under `if (lane == 0)`)
    if (!kPair || crank == 0) {  // pair mode: the leader CTA issues for both
      constexpr uint32_t idesc = make_idesc_bf16(kPair ? 2 * kBM : kBM, BN, 0, kBMn ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int it = 0; it < my_iters; ++it) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_desc_kmajor(smem_u32(smem_a + stage * (kBM * kBK * 2)));
          const uint32_t sb_addr = smem_u32(smem_b + stage * (BN * kBK * 2));
          const uint64_t db = kBMn ? make_desc_mnmajor(sb_addr, 64 * 128) : make_desc_kmajor(sb_addr);
          const uint32_t acc0 = kb ? 1u : 0u;
          if (elect_one()) {
            // K-major: +32 bytes per 16-element K step inside the swizzle atom (encoded >> 4);
            // MN-major: 16 reduction rows = 2048 bytes
            if (kPair) {
              umma_ss_pair(d_tmem, da, db, idesc, acc0);
#pragma unroll
              for (int k = 1; k < kBK / 16; ++k) umma_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, 1u);
              umma_commit_pair_mc(&empty_bar[stage], kMcMask);
              if (kb == num_kb - 1) umma_commit_pair_mc(&tmem_full[as], kMcMask);
            } else {
              umma_ss(d_tmem, da, db, idesc, acc0);
#pragma unroll
              for (int k = 1; k < kBK / 16; ++k) umma_ss(d_tmem, da + 2 * k, db + (kBMn ? 128 : 2) * k, idesc, 1u);
              if (CM > 1) umma_commit_mc(&empty_bar[stage], kMcMask);
              else umma_commit(&empty_bar[stage]);
              if (kb == num_kb - 1) umma_commit(&tmem_full[as]);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4 && ET::kTma) {
    // ---------------- epilogue: TMEM -> registers -> swizzled smem -> TMA store ----------------
    // Every epilogue warp works on its own: the two warps of a TMEM lane quadrant take alternate column chunks of
    // the tile (all CW columns of their 32 rows), stage them in a private [32][128 B] buffer and issue their own
    // TMA store / input load.  No cross-warp barrier: while one warp of a sub-partition waits on TMEM, TMA or its
    // store buffer, the other keeps the FMA / MUFU pipes busy with its GELU.
    constexpr int CW = ET::kCW;
    constexpr int kChunks = BN / CW;
    constexpr int kWarpChunkBytes = 32 * 128;  // 4 KB
    const int q = warp & 3;  // TMEM lane quadrant owned by this warp
    const int half = (warp - 4) >> 2;
    const int ew = warp - 4;
    uint8_t* my_out = stage_out + ew * kWarpChunkBytes;
    uint8_t* my_out2 = stage_out2 + ew * kWarpChunkBytes;
    uint8_t* my_in = stage_in + ew * kWarpChunkBytes;
    uint64_t* my_in_full = &in_full[ew];
    const uint32_t s_out = smem_u32(my_out), s_out2 = smem_u32(my_out2), s_in = smem_u32(my_in);
    const int total_chunks = my_iters * kChunks;
    auto chunk_coords = [&](int n, int& m0, int& col0) {
      int n0;
      tile_coords(n / kChunks, m0, n0);
      m0 += q * 32;
      col0 = n0 + (n % kChunks) * CW;
    };
    if (ET::kHasIn && lane == 0 && half < total_chunks) {  // prefetch the input of this warp's first chunk
      int m0, col0;
      chunk_coords(half, m0, col0);
      mbar_arrive_expect_tx(my_in_full, kWarpChunkBytes);
      tma_load_2d(my_in, &map_in, my_in_full, col0, m0);
    }
    int as = 0;
    uint32_t aphase = 0, in_phase = 0;
    for (int it = 0; it < my_iters; ++it) {
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      // last chunk of this tile that belongs to this warp (-1: none, e.g. single-chunk tiles on the other parity)
      int c_last = kChunks - 1;
      if (((it * kChunks + c_last) & 1) != half) --c_last;
      if (c_last < 0) {
        __syncwarp();
        if (lane == 0) {
          if (kPair && crank != 0) mbar_arrive_remote(&tmem_empty[as], 0);
          else mbar_arrive(&tmem_empty[as]);
        }
      }
      for (int c = 0; c < kChunks; ++c) {
        const int n = it * kChunks + c;
        if ((n & 1) != half) continue;
        int m0, col0;
        chunk_coords(n, m0, col0);
        // accumulator chunk -> registers
        uint32_t v[CW];
        {
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c * CW;
          tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          if constexpr (CW == 64) tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
        }
        if (c == c_last) {  // this warp has read its share of the accumulator tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kPair && crank != 0) mbar_arrive_remote(&tmem_empty[as], 0);  // the leader's MMA warp waits for both CTAs
            else mbar_arrive(&tmem_empty[as]);
          }
        }
        // the staging buffer must have been read out by this warp's previous TMA store
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (ET::kHasIn) mbar_wait(my_in_full, in_phase);
        __syncwarp();
        const float* bias = (EPI != EPI_DGELU && p.bias != nullptr) ? p.bias + col0 : nullptr;
        if (EPI == EPI_BIAS || EPI == EPI_BIAS_GELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(v[8 * j + i]);
            if (bias != nullptr) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + 8 * j));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + 8 * j + 4));
              x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
              x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
            }
            st_shared_v4(s_out + sw128_offset(lane, j), pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]),
                         pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
            if (EPI == EPI_BIAS_GELU)
              st_shared_v4(s_out2 + sw128_offset(lane, j), pack_bf16(gelu_exact(x[0]), gelu_exact(x[1])),
                           pack_bf16(gelu_exact(x[2]), gelu_exact(x[3])), pack_bf16(gelu_exact(x[4]), gelu_exact(x[5])),
                           pack_bf16(gelu_exact(x[6]), gelu_exact(x[7])));
          }
        } else if (EPI == EPI_DGELU) {
          uint32_t hr[32];  // pre-activations of this row: 64 bf16
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 hraw = ld_shared_f4(s_in + sw128_offset(lane, j));
            hr[4 * j] = __float_as_uint(hraw.x); hr[4 * j + 1] = __float_as_uint(hraw.y);
            hr[4 * j + 2] = __float_as_uint(hraw.z); hr[4 * j + 3] = __float_as_uint(hraw.w);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 h0 = unpack_bf16(hr[4 * j]), h1 = unpack_bf16(hr[4 * j + 1]), h2 = unpack_bf16(hr[4 * j + 2]),
                         h3 = unpack_bf16(hr[4 * j + 3]);
            float y[8];
            y[0] = __uint_as_float(v[8 * j + 0]) * gelu_exact_grad(h0.x);
            y[1] = __uint_as_float(v[8 * j + 1]) * gelu_exact_grad(h0.y);
            y[2] = __uint_as_float(v[8 * j + 2]) * gelu_exact_grad(h1.x);
            y[3] = __uint_as_float(v[8 * j + 3]) * gelu_exact_grad(h1.y);
            y[4] = __uint_as_float(v[8 * j + 4]) * gelu_exact_grad(h2.x);
            y[5] = __uint_as_float(v[8 * j + 5]) * gelu_exact_grad(h2.y);
            y[6] = __uint_as_float(v[8 * j + 6]) * gelu_exact_grad(h3.x);
            y[7] = __uint_as_float(v[8 * j + 7]) * gelu_exact_grad(h3.y);
            st_shared_v4(s_out + sw128_offset(lane, j), pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]),
                         pack_bf16(y[6], y[7]));
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 * j + i] = __float_as_uint(y[i]);
          }
          if (p.delta != nullptr) {
            // column sums of this warp's 32 x 64 block of dH (= its share of the fc1 bias gradient): transposing
            // butterfly, 62 shuffles, after which lane l owns columns 2l and 2l+1 (rows beyond M are exact zeros)
#pragma unroll
            for (int bit = 16, n = 32; bit >= 1; bit >>= 1, n >>= 1) {
              const bool up = (lane & bit) != 0;
#pragma unroll
              for (int i = 0; i < n; ++i) {
                const float send = __uint_as_float(up ? v[i] : v[i + n]);
                const float keep = __uint_as_float(up ? v[i + n] : v[i]);
                v[i] = __float_as_uint(keep + __shfl_xor_sync(0xffffffffu, send, bit));
              }
            }
            atomicAdd(p.delta + col0 + 2 * lane, __uint_as_float(v[0]));
            atomicAdd(p.delta + col0 + 2 * lane + 1, __uint_as_float(v[1]));
          }
        } else if (EPI == EPI_DELTA) {
          float dsum = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 oraw = ld_shared_f4(s_in + sw128_offset(lane, j));
            const uint32_t ow[4] = {__float_as_uint(oraw.x), __float_as_uint(oraw.y), __float_as_uint(oraw.z),
                                    __float_as_uint(oraw.w)};
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              pk[i] = pack_bf16(__uint_as_float(v[8 * j + 2 * i]), __uint_as_float(v[8 * j + 2 * i + 1]));
              const float2 g2 = unpack_bf16(pk[i]), o2 = unpack_bf16(ow[i]);  // the rounded dO the attention kernel reads
              dsum = fmaf(g2.x, o2.x, dsum);
              dsum = fmaf(g2.y, o2.y, dsum);
            }
            st_shared_v4(s_out + sw128_offset(lane, j), pk[0], pk[1], pk[2], pk[3]);
          }
          const int m = m0 + lane;
          if (m < p.M) {
            const int bi = m / p.seq_L;
            const int qi = m - bi * p.seq_L;
            p.delta[(static_cast<size_t>(bi) * (p.N >> 6) + (col0 >> 6)) * p.seq_Lp + qi] = dsum;
          }
        } else {  // fp32 outputs: EPI_BIAS_RESID / EPI_F32 (32 columns per chunk)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3]));
            if (bias != nullptr) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + 4 * j));
              o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            }
            if (EPI == EPI_BIAS_RESID) {
              const float4 rs = ld_shared_f4(s_in + sw128_offset(lane, j));
              o.x += rs.x; o.y += rs.y; o.z += rs.z; o.w += rs.w;
            }
            st_shared_v4(s_out + sw128_offset(lane, j), __float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z),
                         __float_as_uint(o.w));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (ET::kHasIn) {  // the input buffer has been consumed: fetch this warp's next chunk
            const int n1 = n + 2;
            if (n1 < total_chunks) {
              int m1, c1;
              chunk_coords(n1, m1, c1);
              mbar_arrive_expect_tx(my_in_full, kWarpChunkBytes);
              tma_load_2d(my_in, &map_in, my_in_full, c1, m1);
            }
          }
          tma_store_2d(&map_out, my_out, col0, m0);
          if (EPI == EPI_BIAS_GELU) tma_store_2d(&map_out2, my_out2, col0, m0);
          tma_store_commit();
        }
        in_phase ^= 1;
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait0();
  } else if (warp >= 8) {
    // direct-store epilogue uses warps 4-7 only; these warps just keep the tmem_empty arrival count
    int as = 0;
    uint32_t aphase = 0;
    for (int it = 0; it < my_iters; ++it) {
      mbar_wait(&tmem_full[as], aphase);  // same cadence as the reading warps
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------- epilogue with direct global stores (EPI_EMBED row remap) -------------------
    const int q = warp & 3;  // TMEM lane quadrant owned by this warp
    int as = 0;
    uint32_t aphase = 0;
    for (int it = 0; it < my_iters; ++it) {
      int m0, n0;
      tile_coords(it, m0, n0);
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      size_t row_off = static_cast<size_t>(row) * p.ldo;
      size_t add_off = 0;
      if (EPI == EPI_EMBED) {
        const int bi = row / p.map_T, t = row - bi * p.map_T;
        row_off = (static_cast<size_t>(bi) * p.map_L + 1 + t) * p.ldo;
        add_off = static_cast<size_t>(t) * p.ldo;
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c * 32, r);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        if (EPI != EPI_DGELU && EPI != EPI_EMBED && p.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
        }
        if (row_ok) {
          if (EPI == EPI_BIAS || EPI == EPI_BIAS_GELU) {
            uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + row_off + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 o;
              o.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
              o.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
              o.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
              o.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
              dst[i] = o;
            }
            if (EPI == EPI_BIAS_GELU) {
              uint4* dst2 = reinterpret_cast<uint4*>(p.out2_bf16 + row_off + col0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 o;
                o.x = pack_bf16(gelu_exact(v[8 * i + 0]), gelu_exact(v[8 * i + 1]));
                o.y = pack_bf16(gelu_exact(v[8 * i + 2]), gelu_exact(v[8 * i + 3]));
                o.z = pack_bf16(gelu_exact(v[8 * i + 4]), gelu_exact(v[8 * i + 5]));
                o.w = pack_bf16(gelu_exact(v[8 * i + 6]), gelu_exact(v[8 * i + 7]));
                dst2[i] = o;
              }
            }
          } else if (EPI == EPI_DGELU) {
            const uint4* hsrc = reinterpret_cast<const uint4*>(p.aux + row_off + col0);
            uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + row_off + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 h4 = __ldg(hsrc + i);
              const float2 h0 = unpack_bf16(h4.x), h1 = unpack_bf16(h4.y), h2 = unpack_bf16(h4.z),
                           h3 = unpack_bf16(h4.w);
              uint4 o;
              o.x = pack_bf16(v[8 * i + 0] * gelu_exact_grad(h0.x), v[8 * i + 1] * gelu_exact_grad(h0.y));
              o.y = pack_bf16(v[8 * i + 2] * gelu_exact_grad(h1.x), v[8 * i + 3] * gelu_exact_grad(h1.y));
              o.z = pack_bf16(v[8 * i + 4] * gelu_exact_grad(h2.x), v[8 * i + 5] * gelu_exact_grad(h2.y));
              o.w = pack_bf16(v[8 * i + 6] * gelu_exact_grad(h3.x), v[8 * i + 7] * gelu_exact_grad(h3.y));
              dst[i] = o;
            }
          } else {  // EPI_BIAS_RESID / EPI_F32 / EPI_EMBED
            float4* dst = reinterpret_cast<float4*>(p.out_f32 + row_off + col0);
            if (EPI == EPI_BIAS_RESID || EPI == EPI_EMBED) {
              const float4* rs = EPI == EPI_EMBED ? reinterpret_cast<const float4*>(p.addend + add_off + col0)
                                                  : reinterpret_cast<const float4*>(p.resid + row_off + col0);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 r4 = rs[i];
                v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w;
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  if (CM > 1) cluster_sync_all();  // no CTA may exit while a peer can still multicast into it / arrive on its barriers
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// gemm_tn: C[Nout, Kout] (+)= sum_m A[m, Nout]^T * B[m, Kout]
// One CTA per (128 x BN output tile, split of the M reduction).
// ---------------------------------------------------------------------------
struct GemmTnParams {
  int M, Nout, Kout;
  int ldc;
  int splits;
  int use_atomic;  // 1: atomicAdd into C (C pre-initialised), 0: plain store (splits must be 1)
  float* C;
  uint32_t lbo_a, lbo_b, sbo;  // descriptor strides (debug-overridable)
};

template <int BN>
struct TnCfg {
  static constexpr int kABytes = kBM * kBK * 2;  // 2 MN chunks of [64 m-rows][64] bf16
  static constexpr int kBBytes = BN * kBK * 2;   // BN/64 chunks
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = 2 * kChunkBytes;  // fp32 [128][32] x 2, SWIZZLE_128B
  static constexpr int kStages = (BN <= 128) ? 6 : 4;
  static constexpr int kTmemCols = (BN <= 128) ? 128 : 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const GemmTnParams p) {
  using Cfg = TnCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* stage_c = smem + kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int k_tiles = p.Kout / BN;
  const int n_tiles = (p.Nout + kBM - 1) / kBM;
  const int tile = blockIdx.x % (k_tiles * n_tiles);
  const int split = blockIdx.x / (k_tiles * n_tiles);
  const int n0 = (tile / k_tiles) * kBM;  // rows of C
  const int k0 = (tile % k_tiles) * BN;   // cols of C
  const int total_mb = (p.M + kBK - 1) / kBK;
  const int mb_per = (total_mb + p.splits - 1) / p.splits;
  const int mb_begin = split * mb_per;
  const int mb_end = min(total_mb, mb_begin + mb_per);
  const int num_mb = max(0, mb_end - mb_begin);
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_wait();

  if (num_mb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int mb = mb_begin; mb < mb_end; ++mb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          uint8_t* sa = smem_a + stage * Cfg::kABytes;
          uint8_t* sb = smem_b + stage * Cfg::kBBytes;
#pragma unroll
          for (int c = 0; c < kBM / 64; ++c)
            tma_load_2d(sa + c * (64 * 128), &map_a, &full_bar[stage], n0 + c * 64, mb * kBK);
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            tma_load_2d(sb + c * (64 * 128), &map_b, &full_bar[stage], k0 + c * 64, mb * kBK);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      {
        constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 1, 1);
        int stage = 0;
        uint32_t phase = 0;
        for (int mb = 0; mb < num_mb; ++mb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_desc_sw128(smem_u32(smem_a + stage * Cfg::kABytes), p.lbo_a, p.sbo);
          const uint64_t db = make_desc_sw128(smem_u32(smem_b + stage * Cfg::kBBytes), p.lbo_b, p.sbo);
          const uint32_t acc0 = mb ? 1u : 0u;
          if (elect_one()) {  // one elected lane inside warp-uniform control flow (no UTCHMMA waterfall)
            // 16 reduction rows = 2 swizzle atoms = 2048 bytes (encoded >> 4 = 128)
            umma_ss(tmem_base, da, db, idesc, acc0);
#pragma unroll
            for (int k = 1; k < kBK / 16; ++k) umma_ss(tmem_base, da + 128 * k, db + 128 * k, idesc, 1u);
            umma_commit(&empty_bar[stage]);
            if (mb == num_mb - 1) umma_commit(tmem_full);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp >= 4) {
      // epilogue: 32-column fp32 chunks -> swizzled smem -> TMA reduce-add (split-K) or TMA store
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const int tid_e = threadIdx.x - 128;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        const int buf = c & 1;
        if (tid_e == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 3, 128;" ::: "memory");
        const uint32_t s_out = smem_u32(stage_c + buf * kChunkBytes);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(s_out + sw128_offset(r, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async_smem();
        asm volatile("bar.sync 3, 128;" ::: "memory");
        if (tid_e == 0) {
          if (p.use_atomic) tma_reduce_add_2d(&map_c, stage_c + buf * kChunkBytes, k0 + c * 32, n0);
          else tma_store_2d(&map_c, stage_c + buf * kChunkBytes, k0 + c * 32, n0);
          tma_store_commit();
        }
      }
      if (tid_e == 0) tma_store_wait0();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
static int g_tn_lbo = 0, g_tn_sbo = 0;  // debug overrides (0 = default)

struct EpiMaps {
  CUtensorMap out, out2, in;
};

template <int BN, int EPI, bool kBMn, int MODE>
static int launch_nt(const CUtensorMap& ma, const CUtensorMap& mb, const GemmNtParams& p, cudaStream_t st) {
  using Cfg = NtCfg<BN, EPI>;
  using ET = EpiTraits<EPI>;
  constexpr int CM = MODE == 0 ? 1 : 2;
  auto kern = gemm_nt_kernel<BN, EPI, kBMn, MODE>;
  EpiMaps em;
  em.out = em.out2 = em.in = ma;  // placeholders for the maps an epilogue does not use
  if (ET::kTma) {
    const uint64_t esz = ET::kF32Out ? 4 : 2;
    void* outp = ET::kF32Out ? static_cast<void*>(p.out_f32) : static_cast<void*>(p.out_bf16);
    if (int e = make_tmap_2d(&em.out, outp, ET::kF32Out, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ldo * esz, ET::kCW, 32))
      return e;
    if (ET::kTwoOut)
      if (int e = make_tmap_2d(&em.out2, p.out2_bf16, false, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ldo * 2, ET::kCW, 32))
        return e;
    if (EPI == EPI_BIAS_RESID)
      if (int e = make_tmap_2d(&em.in, p.resid, true, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ldin * 4, ET::kCW, 32))
        return e;
    if (EPI == EPI_DGELU || EPI == EPI_DELTA)
      if (int e = make_tmap_2d(&em.in, p.aux, false, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ldin * 2, ET::kCW, 32))
        return e;
  }
  DCV_TRY_SMEM_ATTR(kern, Cfg::kSmemBytes);
  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int num_ct = ((m_tiles + CM - 1) / CM) * (p.N / BN);
  const int max_clusters = num_sms() / CM;
  const int clusters = num_ct < max_clusters ? num_ct : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CM);
  cfg.blockDim = dim3(kNtThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CM > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CM;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DCV_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, em.out, em.out2, em.in, p));
  count_launch();
  return 0;
}

template <int BN, int MODE>
static int dispatch_nt_epi(int epi, bool b_mn, const CUtensorMap& ma, const CUtensorMap& mb, const GemmNtParams& p,
                           cudaStream_t st) {
  if (b_mn) {  // dgrad flavours only
    if constexpr (MODE != 2) {
      switch (epi) {
        case EPI_BIAS: return launch_nt<BN, EPI_BIAS, true, MODE>(ma, mb, p, st);
        case EPI_DGELU: return launch_nt<BN, EPI_DGELU, true, MODE>(ma, mb, p, st);
        case EPI_DELTA: return launch_nt<BN, EPI_DELTA, true, MODE>(ma, mb, p, st);
        case EPI_F32: return launch_nt<BN, EPI_F32, true, MODE>(ma, mb, p, st);
      }
    }
    return set_error(DCV_ERR_UNSUPPORTED, "gemm_nn: epilogue %d / mode %d not instantiated", epi, MODE);
  }
  switch (epi) {
    case EPI_BIAS: return launch_nt<BN, EPI_BIAS, false, MODE>(ma, mb, p, st);
    case EPI_BIAS_GELU: return launch_nt<BN, EPI_BIAS_GELU, false, MODE>(ma, mb, p, st);
    case EPI_BIAS_RESID: return launch_nt<BN, EPI_BIAS_RESID, false, MODE>(ma, mb, p, st);
    case EPI_DGELU: return launch_nt<BN, EPI_DGELU, false, MODE>(ma, mb, p, st);
    case EPI_F32: return launch_nt<BN, EPI_F32, false, MODE>(ma, mb, p, st);
    case EPI_EMBED:
      if constexpr (MODE == 0) return launch_nt<BN, EPI_EMBED, false, 0>(ma, mb, p, st);
      break;
  }
  return set_error(DCV_ERR_INVALID, "gemm_nt: unknown epilogue %d", epi);
}

// Cluster size along M for the multicast B operand.  Measured on B200 at the ViT-S shapes (K or N = 384): no gain
// (qkv 50.2 us without vs 51.8 us with a 2-CTA cluster) -- these GEMMs sit at the HBM ridge, not at the L2 -> SM
// limit -- so the default stays 1; dcv_debug_set_nt_cluster(2) switches the multicast path on.
static int g_nt_cluster = 1;
// 1: K-major-B GEMMs (forward Linear layers) run as cta_group::2 pair MMAs (256 x BN tile per 2-CTA cluster).  Correct
// on B200, but measured slower at the ViT-S shapes (qkv 61.7 vs 50.7 us, fc1+GELU 126 vs 108 us): with K = 384 the
// main loop is bounded by shared-memory bandwidth (TMA writes + UMMA operand reads ~ 208 B/clk wanted vs 128 B/clk),
// which the pair mode reduces by only ~15 % while adding cross-CTA barrier latency.  Off by default
// (dcv_debug_set_nt_cluster(3) enables it).
static int g_nt_pair = 0;

// b_mn = false: B is [N][K] (K contiguous);  b_mn = true: B is [K][N] (N contiguous)
int gemm_nt(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epi, const float* bias,
            void* out, void* out2, const float* resid, const void* aux, int ldo, bool b_mn, cudaStream_t st,
            int map_T, int map_L, const float* addend, int ldin, float* delta, int seq_L) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(DCV_ERR_INVALID, "gemm_nt: empty problem %dx%dx%d", M, N, K);
  if (K % 8 || lda % 8 || ldb % 8 || ldo % 8)
    return set_error(DCV_ERR_UNSUPPORTED, "gemm_nt: K/lda/ldb/ldo must be multiples of 8 (16-byte rows)");
  int bn = 0;
  // 128 x 256 tiles for wide forward layers (fc1: 91 -> 84.5 us; operand traffic per MMA flop drops 10 %); the dgrad
  // flavour (MN-major B) measured no gain
  if (!b_mn && N % 256 == 0 && N >= 1024 && epi != EPI_EMBED && !g_nt_pair && g_nt_cluster == 1) bn = 256;
  else if (N % 192 == 0) bn = 192;
  else if (N % 128 == 0) bn = 128;
  else if (N % 64 == 0) bn = 64;
  else return set_error(DCV_ERR_UNSUPPORTED, "gemm_nt: N=%d must be a multiple of 64", N);
  ProfScope prof(b_mn ? PT_GEMM_NN : (epi == EPI_EMBED ? PT_EMBED_GEMM : PT_GEMM_NT), st);
  const bool pair = g_nt_pair && !b_mn && epi != EPI_EMBED && M > kBM && N % 128 == 0;
  const int cm = pair ? 2 : ((epi == EPI_EMBED || M <= kBM) ? 1 : g_nt_cluster);
  CUtensorMap ma, mb;
  if (int e = make_tmap_bf16_2d(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, kBK, kBM)) return e;
  // B boxes are 1/cm of the tile: each CTA of a cluster fetches one slice and multicasts it
  if (b_mn) {
    if (int e = make_tmap_bf16_2d(&mb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, kBK / cm)) return e;
  } else {
    if (int e = make_tmap_bf16_2d(&mb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, kBK, bn / cm)) return e;
  }
  GemmNtParams p;
  p.M = M; p.N = N; p.K = K; p.ldo = ldo; p.bias = bias;
  p.ldin = ldin > 0 ? ldin : ldo;
  if (p.ldin % 8) return set_error(DCV_ERR_UNSUPPORTED, "gemm_nt: ldin must be a multiple of 8");
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out);
  p.out2_bf16 = reinterpret_cast<__nv_bfloat16*>(out2);
  p.out_f32 = reinterpret_cast<float*>(out);
  p.resid = resid;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(aux);
  p.map_T = map_T; p.map_L = map_L; p.addend = addend;
  p.delta = delta; p.seq_L = seq_L; p.seq_Lp = (seq_L + 127) / 128 * 128;
  if (epi == EPI_DELTA && (!b_mn || !aux || !delta || seq_L <= 0 || M % seq_L))
    return set_error(DCV_ERR_INVALID, "gemm_nn: EPI_DELTA needs aux (O), delta, seq_L > 0 dividing M");
  if (epi == EPI_EMBED && (!addend || map_T <= 0 || map_L <= map_T || M % map_T))
    return set_error(DCV_ERR_INVALID, "gemm_nt: EPI_EMBED needs addend, T>0, L>T, M %% T == 0");
  if ((epi == EPI_BIAS_GELU && !out2) || (epi == EPI_BIAS_RESID && !resid) || ((epi == EPI_DGELU || epi == EPI_DELTA) && !aux) || !out)
    return set_error(DCV_ERR_INVALID, "gemm_nt: missing buffer for epilogue %d", epi);
  if (pair) {
    switch (bn) {
      case 192: return dispatch_nt_epi<192, 2>(epi, b_mn, ma, mb, p, st);
      default: return dispatch_nt_epi<128, 2>(epi, b_mn, ma, mb, p, st);
    }
  }
  if (cm == 2) {
    switch (bn) {
      case 192: return dispatch_nt_epi<192, 1>(epi, b_mn, ma, mb, p, st);
      case 128: return dispatch_nt_epi<128, 1>(epi, b_mn, ma, mb, p, st);
      default: return dispatch_nt_epi<64, 1>(epi, b_mn, ma, mb, p, st);
    }
  }
  switch (bn) {
    case 256: return dispatch_nt_epi<256, 0>(epi, b_mn, ma, mb, p, st);
    case 192: return dispatch_nt_epi<192, 0>(epi, b_mn, ma, mb, p, st);
    case 128: return dispatch_nt_epi<128, 0>(epi, b_mn, ma, mb, p, st);
    default: return dispatch_nt_epi<64, 0>(epi, b_mn, ma, mb, p, st);
  }
}

// cm: 1 = single-CTA tiles, 2 = B multicast in 2-CTA clusters, 3 = cta_group::2 pair MMA for K-major B (default)
void debug_set_nt_cluster(int cm) {
  g_nt_cluster = (cm == 2) ? 2 : 1;
  g_nt_pair = (cm == 3) ? 1 : 0;
}

template <int BN>
static int launch_tn(const CUtensorMap& ma, const CUtensorMap& mb, GemmTnParams p, cudaStream_t st) {
  using Cfg = TnCfg<BN>;
  auto kern = gemm_tn_kernel<BN>;
  DCV_TRY_SMEM_ATTR(kern, Cfg::kSmemBytes);
  const int tiles = ((p.Nout + kBM - 1) / kBM) * (p.Kout / BN);
  const int total_mb = (p.M + kBK - 1) / kBK;
  int splits = p.splits;
  if (splits <= 0) {  // fill the machine once
    splits = num_sms() / tiles;
    if (splits < 1) splits = 1;
    if (splits > total_mb) splits = total_mb;
  }
  if (!p.use_atomic) splits = 1;
  p.splits = splits;
  p.lbo_a = g_tn_lbo ? g_tn_lbo : 64 * 128;
  p.lbo_b = g_tn_lbo ? g_tn_lbo : 64 * 128;
  p.sbo = g_tn_sbo ? g_tn_sbo : 1024;
  CUtensorMap mc;
  if (int e = make_tmap_2d(&mc, p.C, true, (uint64_t)p.Kout, (uint64_t)p.Nout, (uint64_t)p.ldc * 4, 32, kBM)) return e;
  DCV_CUDA(launch_pdl(kern, dim3(tiles * splits), dim3(256), Cfg::kSmemBytes, st, ma, mb, mc, p));
  count_launch();
  return 0;
}

int gemm_tn(const void* A, int lda, const void* B, int ldb, int M, int Nout, int Kout, float* C, int ldc,
            int accumulate, int splits, cudaStream_t st) {
  if (M <= 0 || Nout <= 0 || Kout <= 0) return set_error(DCV_ERR_INVALID, "gemm_tn: empty problem");
  if (lda % 8 || ldb % 8 || Nout % 64 || ldc % 4)
    return set_error(DCV_ERR_UNSUPPORTED, "gemm_tn: lda/ldb %% 8, Nout %% 64, ldc %% 4 required");
  int bn = 0;
  if (Kout % 192 == 0) bn = 192;
  else if (Kout % 128 == 0) bn = 128;
  else if (Kout % 64 == 0) bn = 64;
  else return set_error(DCV_ERR_UNSUPPORTED, "gemm_tn: Kout=%d must be a multiple of 64", Kout);
  ProfScope prof(PT_GEMM_TN, st);
  CUtensorMap ma, mb;
  if (int e = make_tmap_bf16_2d(&ma, A, (uint64_t)Nout, (uint64_t)M, (uint64_t)lda * 2, 64, kBK)) return e;
  if (int e = make_tmap_bf16_2d(&mb, B, (uint64_t)Kout, (uint64_t)M, (uint64_t)ldb * 2, 64, kBK)) return e;
  GemmTnParams p;
  p.M = M; p.Nout = Nout; p.Kout = Kout; p.ldc = ldc; p.splits = splits; p.use_atomic = accumulate; p.C = C;
  switch (bn) {
    case 192: return launch_tn<192>(ma, mb, p, st);
    case 128: return launch_tn<128>(ma, mb, p, st);
    default: return launch_tn<64>(ma, mb, p, st);
  }
}

void debug_set_tn_desc(int lbo, int sbo) { g_tn_lbo = lbo; g_tn_sbo = sbo; }

}  // namespace dcv
