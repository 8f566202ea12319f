#include <cuda_runtime.h>
#include "../../diverse_channel_vit_b200/csrc/common.cuh"
using namespace dcv;
template <int VAR>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = slot;
  if (VAR == 2) tm = __shfl_sync(0xffffffffu, tm, 0);
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
  if (warp == 1) {
    if (VAR == 0) {
      if (lane == 0) {
        for (int it = 0; it < iters; ++it) {
          const uint64_t db = make_desc_mnmajor(smem_u32(smem + (it & 1) * 16384), 8192);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ts(tm + 256, tm + 448 + 8 * kk, db + 128 * kk, idesc, 1u);
        }
      }
    } else {
      for (int it = 0; it < iters; ++it) {
        const uint64_t db = make_desc_mnmajor(smem_u32(smem + (it & 1) * 16384), 8192);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ts(tm + 256, tm + 448 + 8 * kk, db + 128 * kk, idesc, 1u);
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}
template __global__ void k<0>(long long*, int);
template __global__ void k<1>(long long*, int);
template __global__ void k<2>(long long*, int);
