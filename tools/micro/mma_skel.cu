// Microbenchmark: the barrier protocol of the attention-backward kernel (V == 3 schedule) with no arithmetic at all --
// two MMA issuer warps, 16 "compute" warps that only wait / arrive, static operands.  How many clocks per phase does
// the hand-shake structure itself cost, against the ~1650 clk the same 32 MMAs take when issued free-running
// (mma_seq2)?  Variants (argv): bit0 = compute warps do not wait for s_full (S pre-load dropped from the chain),
// bit1 = compute warps do not wait for p_free, bit2 = compute warps do not wait for dp_full,
// bit3 = only ONE lane per compute warp polls / arrives (barrier counts 16 instead of 512).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../diverse_channel_vit_b200/csrc/common.cuh"
using namespace dcv;

constexpr int kThreads = 32 * 20;

struct Bars {
  uint64_t s_full, dp_full, dp_consumed, phase_done, p_free, done;
};

__global__ void __launch_bounds__(kThreads, 1) skel_kernel(long long* out, int iters, int feat) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ Bars bars;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (i + 1) * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x807f807fu) | 0x3f003f00u;
  }
  const bool one_lane = feat & 8;
  const int cnt = one_lane ? 16 : 512;
  if (threadIdx.x == 0) {
    mbar_init(&bars.s_full, 1); mbar_init(&bars.dp_full, 1); mbar_init(&bars.p_free, 1); mbar_init(&bars.done, 1);
    mbar_init(&bars.dp_consumed, cnt); mbar_init(&bars.phase_done, cnt);
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0), id_kv = make_idesc_bf16(128, 64, 0, 1),
                     id_dq = make_idesc_bf16(128, 64, 1, 1);
  const uint32_t tS = tm, tdP = tm + 128, tdV = tm + 256, tdK = tm + 320, tdQ = tm + 384, tP = tm + 448;
  uint8_t* sK = smem; uint8_t* sV = smem + 16384; uint8_t* sQ = smem + 32768; uint8_t* sdO = sQ + 4 * 16384;
  uint8_t* sdS = sdO + 2 * 16384;
  const int n_q = iters;

  if (warp < 16) {  // "compute": phase i = 0..n_q
    const bool act = !one_lane || lane == 0;
    for (int i = 0; i <= n_q; ++i) {
      if (act) {
        if (i > 0) {
          if (!(feat & 4)) mbar_wait(&bars.dp_full, (i - 1) & 1);
          tc_fence_after();
          tc_fence_before();
          mbar_arrive(&bars.dp_consumed);
        }
        if (i < n_q) {
          if (i > 0 && !(feat & 2)) mbar_wait(&bars.p_free, (i - 1) & 1);
          if (i + 1 < n_q && !(feat & 1)) mbar_wait(&bars.s_full, (i + 1) & 1);
          tc_fence_after();
        } else {
          mbar_wait(&bars.p_free, (i - 1) & 1);
        }
        tc_fence_before();
        mbar_arrive(&bars.phase_done);
      }
      __syncwarp();
    }
  } else if (warp == 17) {  // front: S_0, S_1, then dP_i
    const uint64_t dK_k = make_desc_kmajor(smem_u32(sK)), dV_k = make_desc_kmajor(smem_u32(sV));
    for (int i = -1; i < n_q; ++i) {
      if (i < 1 && i + 1 < n_q) {
        const uint64_t dQ_k = make_desc_kmajor(smem_u32(sQ + ((i + 1) & 3) * 16384));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tS, dK_k + 2 * k, dQ_k + 2 * k, id_s, k ? 1u : 0u);
          umma_commit(&bars.s_full);
        }
        __syncwarp();
      }
      if (i >= 0) {
        if (i > 0) mbar_wait(&bars.dp_consumed, (i - 1) & 1);
        tc_fence_after();
        const uint64_t dO_k = make_desc_kmajor(smem_u32(sdO + (i & 1) * 16384));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tdP, dV_k + 2 * k, dO_k + 2 * k, id_s, k ? 1u : 0u);
          umma_commit(&bars.dp_full);
        }
        __syncwarp();
      }
    }
  } else if (warp == 18) {  // back
    const uint64_t dK_k = make_desc_kmajor(smem_u32(sK));
    const uint64_t dK_mn = make_desc_mnmajor(smem_u32(sK), 16384);
    long long t0 = clock64();
    for (int i = 0; i <= n_q; ++i) {
      const int j = i - 1;
      mbar_wait(&bars.phase_done, i & 1);
      tc_fence_after();
      if (i + 2 < n_q) {
        const uint64_t dQ_k = make_desc_kmajor(smem_u32(sQ + ((i + 2) & 3) * 16384));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tS, dK_k + 2 * k, dQ_k + 2 * k, id_s, k ? 1u : 0u);
          umma_commit(&bars.s_full);
        }
        __syncwarp();
      }
      if (i < n_q) {
        const uint64_t dO_mn = make_desc_mnmajor(smem_u32(sdO + (i & 1) * 16384), 16384);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ts(tdV, tP + 8 * k, dO_mn + 128 * k, id_kv, (i | k) ? 1u : 0u);
          umma_commit(&bars.p_free);
        }
        __syncwarp();
      }
      if (i > 0) {
        const uint64_t dQ_mn = make_desc_mnmajor(smem_u32(sQ + (j & 3) * 16384), 16384);
        const uint64_t dS_k0 = make_desc_kmajor(smem_u32(sdS + (j & 1) * 32768)), dS_k1 = dS_k0 + (16384 >> 4);
        const uint64_t dS_mn = make_desc_mnmajor(smem_u32(sdS + (j & 1) * 32768), 16384);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ss(tdK, (k < 4 ? dS_k0 : dS_k1) + 2 * (k & 3), dQ_mn + 128 * k, id_kv, (j | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ss(tdQ, dS_mn + 128 * k, dK_mn + 128 * k, id_dq, k ? 1u : 0u);
        }
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(&bars.done);
    __syncwarp();
    mbar_wait(&bars.done, 0);
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main(int argc, char** argv) {
  long long* d; cudaMalloc(&d, 64);
  const int smem = 200 * 1024 + 2048;
  cudaFuncSetAttribute(skel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 400;
  for (int a = 1; a < argc; ++a) {
    const int f = atoi(argv[a]);
    skel_kernel<<<148, kThreads, smem>>>(d, iters, f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[1]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("skel feat %2d [%s%s%s%s]: %.0f clk/phase  %s\n", f, f & 1 ? "no-s_full-wait " : "", f & 2 ? "no-p_free-wait " : "",
           f & 4 ? "no-dp_full-wait " : "", f & 8 ? "one-lane " : "", (double)h[0] / iters, cudaGetErrorString(e));
    fflush(stdout);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
