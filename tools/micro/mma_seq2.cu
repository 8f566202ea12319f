// Microbenchmark: the per-(q-tile, k-tile) MMA sequence of the attention backward with the kernel's in-situ features
// added one at a time (bit mask): which of them turns the ~1500-1700 clk sequence into the ~4000 clk the kernel's
// timeline shows?
//   bit0  two issuer warps (front: S^T, dP^T; back: dV, dK, dQ) instead of one thread
//   bit1  whole warp walks the loop, elect_one() issues (the kernel's pattern) instead of `if (lane == 0)`
//   bit2  a TMA warp streams Q / dO / stats (33 KB per pair, bulk copies) into rotating stages, paced by the back warp
//   bit3  8 commits per pair (the kernel's count) instead of 4
//   bit4  8 more warps store 64 KB per pair into shared memory (dS^T + dQ staging traffic)
//   bit5  8 more warps read 256 KB per pair out of TMEM (S^T, dP^T, dQ reads of the compute / drain warps)
//   bit7  16 warps poll an mbarrier (try_wait loop) the whole time
//   bit6  back and front are coupled like in the kernel: back's burst of pair i only after front's S of pair i completed
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../diverse_channel_vit_b200/csrc/common.cuh"
using namespace dcv;

constexpr int kThreads = 32 * 20;

__global__ void __launch_bounds__(kThreads, 1) seq2_kernel(long long* out, int iters, int feat, const uint8_t* gsrc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[16];
  __shared__ uint32_t slot;
  __shared__ volatile int stop_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (i + 1) * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x807f807fu) | 0x3f003f00u;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bar[i], i == 14 ? 512 : 1);
    stop_flag = 0;
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const bool two = feat & 1, elect = feat & 2, tma = feat & 4, many_commits = feat & 8, couple = feat & 64;
  constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0), id_kv = make_idesc_bf16(128, 64, 0, 1),
                     id_dq = make_idesc_bf16(128, 64, 1, 1);
  const uint32_t tS = tm, tdP = tm + 128, tdV = tm + 256, tdK = tm + 320, tdQ = tm + 384, tP = tm + 448;
  // smem: K 16K | V 16K | Q x3 48K | dO x2 32K | dS x2 64K | scratch 24K
  uint8_t* sK = smem; uint8_t* sV = smem + 16384; uint8_t* sQ = smem + 32768; uint8_t* sdO = sQ + 3 * 16384;
  uint8_t* sdS = sdO + 2 * 16384; uint8_t* scratch = sdS + 4 * 16384;

  auto front = [&](int it, bool issue) {
    const uint64_t dK_k = make_desc_kmajor(smem_u32(sK)), dV_k = make_desc_kmajor(smem_u32(sV));
    const uint64_t dQ_k = make_desc_kmajor(smem_u32(sQ + (it % 3) * 16384)), dO_k = make_desc_kmajor(smem_u32(sdO + (it & 1) * 16384));
    if (tma) { mbar_wait(&bar[8 + (it % 3)], (it / 3) & 1); tc_fence_after(); }
    if (issue) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tS, dK_k + 2 * k, dQ_k + 2 * k, id_s, k ? 1u : 0u);
      umma_commit(&bar[0]);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tdP, dV_k + 2 * k, dO_k + 2 * k, id_s, k ? 1u : 0u);
      umma_commit(&bar[1]);
      if (many_commits) umma_commit(&bar[2]);
    }
  };
  auto back = [&](int it, bool issue) {
    const uint64_t dQ_mn = make_desc_mnmajor(smem_u32(sQ + (it % 3) * 16384), 16384);
    const uint64_t dO_mn = make_desc_mnmajor(smem_u32(sdO + (it & 1) * 16384), 16384);
    const uint64_t dK_mn = make_desc_mnmajor(smem_u32(sK), 16384);
    const uint64_t dS_k0 = make_desc_kmajor(smem_u32(sdS + (it & 1) * 32768)), dS_k1 = dS_k0 + (16384 >> 4);
    const uint64_t dS_mn = make_desc_mnmajor(smem_u32(sdS + (it & 1) * 32768), 16384);
    if (couple) { mbar_wait(&bar[0], it & 1); tc_fence_after(); }
    if (issue) {
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_ts(tdV, tP + 8 * k, dO_mn + 128 * k, id_kv, (it | k) ? 1u : 0u);
      umma_commit(&bar[3]);
      if (many_commits) umma_commit(&bar[4]);
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_ss(tdK, (k < 4 ? dS_k0 : dS_k1) + 2 * (k & 3), dQ_mn + 128 * k, id_kv, (it | k) ? 1u : 0u);
      umma_commit(&bar[11 + (it % 3)]);  // frees the Q stage for the TMA warp
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_ss(tdQ, dS_mn + 128 * k, dK_mn + 128 * k, id_dq, k ? 1u : 0u);
      umma_commit(&bar[5]);
      if (many_commits) umma_commit(&bar[6]);
    }
  };

  if (warp == 1) {  // back (or the single issuer)
    long long t0 = clock64();
    if (elect) {
      for (int it = 0; it < iters; ++it) {
        if (!two) { front(it, elect_one()); __syncwarp(); }
        back(it, elect_one());
        __syncwarp();
      }
    } else if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        if (!two) front(it, true);
        back(it, true);
      }
    }
    long long t1 = clock64();
    if (lane == 0) {
      umma_commit(&bar[7]);
      mbar_wait(&bar[7], 0);
      long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
      stop_flag = 1;
    }
    __syncwarp();
  } else if (warp == 2 && two) {  // front
    if (elect) {
      for (int it = 0; it < iters; ++it) { front(it, elect_one()); __syncwarp(); }
    } else if (lane == 0) {
      for (int it = 0; it < iters; ++it) front(it, true);
    }
  } else if (warp == 3 && tma) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int st = it % 3;
        if (it >= 3) mbar_wait(&bar[11 + st], ((it / 3) - 1) & 1);
        mbar_arrive_expect_tx(&bar[8 + st], 16384 + 16384 + 1024);
        bulk_load_1d(sQ + st * 16384, gsrc + (size_t)(it % 64) * 40960, 16384, &bar[8 + st]);
        bulk_load_1d(sdO + (it & 1) * 16384, gsrc + (size_t)(it % 64) * 40960 + 16384, 16384, &bar[8 + st]);
        bulk_load_1d(scratch + 16384, gsrc + (size_t)(it % 64) * 40960 + 32768, 1024, &bar[8 + st]);
      }
    }
  } else if (warp >= 4 && (feat & (256 | 512 | 1024))) {
    // bit8: 16 warps execute tcgen05.fence::after/before_thread_sync pairs, bit9: tcgen05.wait::ld / ::st (nothing
    // outstanding), bit10: mbarrier.arrive on a 512-count barrier -- each at the kernel's per-phase rate
    long long next = clock64();
    while (!stop_flag) {
      if (feat & 256) { tc_fence_after(); tc_fence_before(); tc_fence_after(); tc_fence_before(); }
      if (feat & 512) { tmem_ld_wait(); tmem_st_wait(); }
      if (feat & 1024) { mbar_arrive(&bar[14]); mbar_arrive(&bar[14]); mbar_arrive(&bar[14]); }
      next += 1000;
      while (clock64() < next && !stop_flag) {}
    }
  } else if (warp >= 4 && (feat & 128)) {
    // 16 warps polling an mbarrier that does not complete (what the kernel's 20 waiting warps do most of the time)
    uint32_t n = 0;
    while (true) {
      if (mbar_try_wait(&bar[15], 0)) break;
      if ((++n & 63) == 0 && stop_flag) break;
    }
  } else if (warp >= 4 && warp < 12 && (feat & 16)) {
    // shared-memory store traffic: 8 warps x 16 st.shared.v4 x 512 B = 64 KB per ~1600 clk; paced by clock
    const uint32_t base = smem_u32(scratch) + (warp - 4) * 2048 + lane * 16;
    long long next = clock64();
    while (!stop_flag) {
#pragma unroll
      for (int j = 0; j < 4; ++j) st_shared_v4(base + j * 512, lane, j, warp, 7);
      next += 400;  // 4 stores per 400 clk per warp = 16 KB / 1600 clk / warp... x8 warps = 64 KB per 1600 clk
      while (clock64() < next && !stop_flag) {}
    }
  } else if (warp >= 12 && warp < 20 && (feat & 32)) {
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    long long next = clock64();
    uint32_t acc = 0;
    while (!stop_flag) {
      uint32_t r[32];
      tmem_ld32(tm + lane_base + ((warp >> 2) & 1) * 128, r);
      tmem_ld_wait();
      acc += r[lane];
      next += 200;  // 8 warps x 4 KB per 200 clk = 256 KB per 1600 clk
      while (clock64() < next && !stop_flag) {}
    }
    if (acc == 0x12345678u) out[7] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main(int argc, char** argv) {
  long long* d; cudaMalloc(&d, 64);
  uint8_t* g; cudaMalloc(&g, 64 * 40960);
  cudaMemset(g, 0x3c, 64 * 40960);
  const int smem = 200 * 1024 + 2048;
  cudaFuncSetAttribute(seq2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 1000;
  const int all[] = {0, 3, 7, 16, 32, 48, 3 + 16, 3 + 32, 3 + 48, 7 + 16, 7 + 32, 7 + 48};
  int feats[16], nf = 0;
  if (argc > 1) { for (int i = 1; i < argc && nf < 16; ++i) feats[nf++] = atoi(argv[i]); }  // one process per risky mask
  else { for (int f : all) feats[nf++] = f; }
  for (int fi = 0; fi < nf; ++fi) {
    const int f = feats[fi];
    seq2_kernel<<<148, kThreads, smem>>>(d, iters, f, g);
    cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("feat %4d [%s%s%s%s%s%s%s%s%s%s]: issue %.0f, complete %.0f clk/pair  %s\n", f, f & 1 ? "2warps " : "", f & 2 ? "elect " : "",
           f & 4 ? "tma " : "", f & 8 ? "8commits " : "", f & 16 ? "sts " : "", f & 32 ? "tmem-ld " : "", f & 128 ? "16-warps-polling " : "", f & 256 ? "tc-fences " : "", f & 512 ? "tc-waits " : "", f & 1024 ? "arrives " : "",
           (double)h[0] / iters, (double)h[1] / iters, cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
  }
  return 0;
}
