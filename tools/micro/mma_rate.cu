// Microbenchmark: issue rate / latency of tcgen05.mma (kind::f16, cta_group::1, M=128) from one thread,
// operands resident in shared memory (garbage data), accumulating into TMEM.  Prints cycles per MMA.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../diverse_channel_vit_b200/csrc/common.cuh"
using namespace dcv;

// LD: warps 4-7 hammer TMEM with tcgen05.ld (columns 320..447, away from the accumulator) while the MMAs run;
// SM: warps 4-7 hammer shared memory with 16-byte stores instead
template <int N, bool TS, bool BMN, int NOISE>
__global__ void __launch_bounds__(256, 1) mma_rate_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 * 64 * 2 + 256 * 64 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, BMN ? 1 : 0);
    const uint64_t da = make_desc_kmajor(smem_u32(smem));
    const uint32_t sb = smem_u32(smem + 128 * 64 * 2);
    const uint64_t db = BMN ? make_desc_mnmajor(sb, 64 * 128) : make_desc_kmajor(sb);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS) umma_ts(tm, tm + 256 + 8 * k, db + (BMN ? 128 : 2) * k, idesc, 1u);
        else umma_ss(tm, da + 2 * k, db + (BMN ? 128 : 2) * k, idesc, 1u);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    stop = 1;
  } else if (warp >= 4 && NOISE == 1) {
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    while (!stop) {
      uint32_t r[32];
      tmem_ld32(tm + lane_base + 320, r);
      tmem_ld32(tm + lane_base + 352, r);
      tmem_ld_wait();
      acc += r[0];
    }
    if (acc == 0x12345) out[3] = acc;
  } else if (warp >= 4 && NOISE == 2) {
    const uint32_t base = smem_u32(smem + 128 * 64 * 2 + 256 * 64 * 2) + 0;  // scratch beyond the operands (2 KB)
    uint32_t k = 0;
    while (!stop) {
#pragma unroll
      for (int j = 0; j < 8; ++j) st_shared_v4(base + ((threadIdx.x & 127) * 16), k, k, k, k);
      ++k;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, bool TS, bool BMN, int NOISE = 0>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 16);
  const int smem = 128 * 64 * 2 + 256 * 64 * 2 + 4096;
  cudaFuncSetAttribute(mma_rate_kernel<N, TS, BMN, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 256;
  for (int grid : {1, 148}) {
    mma_rate_kernel<N, TS, BMN, NOISE><<<grid, 256, smem>>>(d, iters);
    mma_rate_kernel<N, TS, BMN, NOISE><<<grid, 256, smem>>>(d, iters);
    cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-28s grid %3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d)  err=%s\n", name, grid,
           (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), N / 4 > 32 ? N / 2 : N / 2, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(d);
}

int main() {
  run<64, false, false>("SS N=64  B K-major");
  run<128, false, false>("SS N=128 B K-major");
  run<192, false, false>("SS N=192 B K-major");
  run<256, false, false>("SS N=256 B K-major");
  run<64, false, true>("SS N=64  B MN-major");
  run<192, false, true>("SS N=192 B MN-major");
  run<64, true, true>("TS N=64  B MN-major");
  run<128, true, false>("TS N=128 B K-major");
  run<128, false, false, 1>("SS N=128 + LDTM noise");
  run<64, true, true, 1>("TS N=64  + LDTM noise");
  run<64, false, true, 1>("SS N=64  + LDTM noise");
  run<128, false, false, 2>("SS N=128 + STS noise");
  run<64, false, true, 2>("SS N=64  + STS noise");
  return 0;
}
