// Host-side helpers shared by the launchers: error reporting, device queries,
// TMA tensor-map construction.  Internal C++ declarations of every launcher that
// api.cu exports through the C ABI (include/dcvit.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dcvit.h"

namespace dcv {

int set_error(int code, const char* fmt, ...);
int num_sms();
void count_launch(int n = 1);
// 2-D bf16 row-major tensor [outer][inner], 128-byte-swizzled boxes; OOB reads return zeros.
int make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer);
// 3-D bf16 tensor [d2][d1][d0] with explicit byte strides for d1, d2.
int make_tmap_bf16_3d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2);

#define DCV_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::dcv::set_error(DCV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                                  \
  } while (0)

// ---- gemm.cu ----
int gemm_nt(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epi, const float* bias,
            void* out, void* out2, const float* resid, const void* aux, int ldo, bool b_mn, cudaStream_t st);
int gemm_tn(const void* A, int lda, const void* B, int ldb, int M, int Nout, int Kout, float* C, int ldc,
            int accumulate, int splits, cudaStream_t st);
void debug_set_tn_desc(int lbo, int sbo);

// ---- attention.cu ----
int attn_fwd(const void* qkv, void* o, float* lse2, int B, int L, int H, float scale, cudaStream_t st);
int attn_bwd(const void* qkv, const void* o, const void* dO, const float* lse2, float* delta, float* dq_acc,
             void* dqkv, int B, int L, int H, float scale, cudaStream_t st);

}  // namespace dcv
