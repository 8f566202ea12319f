// Microbenchmark: TMA (cp.async.bulk.tensor 2D, SWIZZLE_128B boxes of [128 rows][64 bf16] = 16 KB) load throughput
// per SM with all SMs active, from an L2-resident or a DRAM-sized source.  Prints bytes / clock / SM.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../diverse_channel_vit_b200/csrc/common.cuh"
using namespace dcv;

__global__ void __launch_bounds__(64, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap map, long long* out,
                                                          int iters, int rows_total, int kblocks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int S = 8;
  __shared__ uint64_t bar[S];
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int m_tiles = rows_total / 128;
    long long t0 = clock64();
    for (int it = 0; it < iters + S; ++it) {
      const int s = it % S;
      if (it >= S) mbar_wait(&bar[s], ((it / S) - 1) & 1);
      if (it < iters) {
        const int tile = (blockIdx.x * 977 + it * 131) % (m_tiles * kblocks);
        mbar_arrive_expect_tx(&bar[s], 16384);
        tma_load_2d(smem + s * 16384, &map, &bar[s], (tile % kblocks) * 64, (tile / kblocks) * 128);
      }
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const int K = 512;  // columns (bf16)
  for (long long rows : {8192LL, 2097152LL}) {  // 8 MB (L2) and 2 GB (DRAM)
    void* d; cudaMalloc(&d, rows * K * 2); cudaMemset(d, 0, rows * K * 2);
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    long long* out; cudaMalloc(&out, 148 * 8);
    const int smem = 8 * 16384 + 2048;
    cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {1, 16, 74, 148}) {
      const int iters = 2000;
      tma_rate_kernel<<<grid, 64, smem>>>(m, out, iters, (int)rows, K / 64);
      tma_rate_kernel<<<grid, 64, smem>>>(m, out, iters, (int)rows, K / 64);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("source %5lld MB grid %3d: %.1f B/clk/SM  (%.0f B/clk chip)  %s\n", rows * K * 2 >> 20, grid,
             16384.0 * iters / mx, 16384.0 * iters / mx * grid, cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(d); cudaFree(out);
  }
  return 0;
}
