// Microbenchmark: per-SM issue rates of the instructions on the attention softmax / dS path:
// MUFU.EX2, F2FP.BF16.PACK_AB, FFMA, and their mixes (are conversions on the MUFU pipe?).
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 1) rate_kernel(float* out, long long* cyc, int iters, float seed) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed + threadIdx.x * 1e-3f + i;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0 || MODE == 3 || MODE == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 2 || MODE == 4) x[i] = fmaf(x[i], 1.0001f, 0.5f);
    }
    if (MODE == 1 || MODE == 3) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i + 1]), "f"(x[i]));
        acc ^= r;
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, double ops_per_iter_per_thread) {
  float* o; long long* c;
  cudaMalloc(&o, 148 * 256 * 4); cudaMalloc(&c, 8);
  const int iters = 4096;
  rate_kernel<MODE><<<148, 256>>>(o, c, iters, 0.25f);
  rate_kernel<MODE><<<148, 256>>>(o, c, iters, 0.25f);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %8.1f clk/iter  -> %6.2f thread-ops/clk/SM  (%s)\n", name, (double)h / iters,
         ops_per_iter_per_thread * 256 * iters / (double)h, cudaGetErrorString(cudaGetLastError()));
  cudaFree(o); cudaFree(c);
}

int main() {
  run<0>("16 ex2", 16);
  run<1>("8 cvt.bf16x2 (+8 xor)", 8);
  run<2>("16 ffma", 16);
  run<3>("16 ex2 + 8 cvt.bf16x2", 24);
  run<4>("16 ex2 + 16 ffma", 32);
  return 0;
}
