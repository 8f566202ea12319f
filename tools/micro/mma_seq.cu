// Microbenchmark: the per-(q-tile, k-tile) MMA sequence of the fused attention backward, issued by one thread with
// static shared-memory operands: S^T(4 x SS N=128) dP^T(4 x SS N=128) dV(8 x TS N=64) dK(8 x TS N=64) dQ(8 x SS N=64,
// A MN-major), each group into its own TMEM accumulator, with/without a commit after every group.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../diverse_channel_vit_b200/csrc/common.cuh"
using namespace dcv;

template <int MODE>
__global__ void __launch_bounds__(128, 1) seq_kernel(long long* out, int iters, int stage_stride, int random_fill) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0x3c003c00u;
    if (random_fill) {  // random signs / mantissas, exponents around 1.0 (what real activations look like to the datapath)
      uint32_t h = (i + 1) * 2654435761u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      v = (h & 0x807f807fu) | 0x3f003f00u | ((h >> 3) & 0x00800080u);
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0), id_kv = make_idesc_bf16(128, 64, 0, 1),
                       id_dq = make_idesc_bf16(128, 64, 1, 1);
    const uint32_t tS = tm, tdP = tm + 128, tdV = tm + 256, tdK = tm + 320, tdQ = tm + 384, tP = tm + 448;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint8_t* sK = smem, *sV = smem + 16384, *sQ = smem + 32768 + (it & 1) * stage_stride, *sdO = sQ + 32768,
               *sdS = smem + 131072 - 32768;
      const uint64_t dK_k = make_desc_kmajor(smem_u32(sK)), dV_k = make_desc_kmajor(smem_u32(sV));
      const uint64_t dQ_k = make_desc_kmajor(smem_u32(sQ)), dO_k = make_desc_kmajor(smem_u32(sdO));
      const uint64_t dQ_mn = make_desc_mnmajor(smem_u32(sQ), 16384), dO_mn = make_desc_mnmajor(smem_u32(sdO), 16384);
      const uint64_t dK_mn = make_desc_mnmajor(smem_u32(sK), 16384), dS_mn = make_desc_mnmajor(smem_u32(sdS), 16384);
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_ts(tdV, tP + 8 * k, dO_mn + 128 * k, id_kv, (it | k) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tS, dK_k + 2 * k, dQ_k + 2 * k, id_s, k ? 1u : 0u);
      if (MODE >= 1) umma_commit(&bar[0]);
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_ts(tdK, tdP + (k < 4 ? 0 : 64) + 8 * (k & 3), dQ_mn + 128 * k, id_kv, (it | k) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_ss(tdQ, dS_mn + 128 * k, dK_mn + 128 * k, id_dq, k ? 1u : 0u);
      if (MODE >= 1) { umma_commit(&bar[1]); umma_commit(&bar[2]); }
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tdP, dV_k + 2 * k, dO_k + 2 * k, id_s, k ? 1u : 0u);
      if (MODE >= 1) umma_commit(&bar[3]);
      if (MODE == 2) {  // wait for completion of everything issued so far (no overlap between pairs)
        umma_commit(&bar[4]);
        mbar_wait(&bar[4], it & 1);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar[5]);
    mbar_wait(&bar[5], 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int MODE>
void run(const char* name, int random_fill = 0) {
  long long* d; cudaMalloc(&d, 64);
  const int smem = 160 * 1024 + 2048;
  cudaFuncSetAttribute(seq_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  seq_kernel<MODE><<<148, 128, smem>>>(d, iters, 0, random_fill);
  seq_kernel<MODE><<<148, 128, smem>>>(d, iters, 0, random_fill);
  cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-44s issue %.0f cyc/pair, complete %.0f cyc/pair (32 MMAs; calibrated sum 1640)  %s\n", name,
         (double)h[0] / iters, (double)h[1] / iters, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<0>("sequence, no commits");
  run<1>("sequence, commit after each group");
  run<2>("sequence, + wait for completion every pair");
  run<0>("RANDOM data: sequence, no commits", 1);
  run<1>("RANDOM data: commit after each group", 1);
  run<2>("RANDOM data: + wait every pair", 1);
  return 0;
}
