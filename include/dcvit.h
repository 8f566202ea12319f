/* dcvit.h -- C ABI of libdcvit.so, the sm_100a kernel library behind the drop-in
 * DiChaViT module (diverse_channel_vit_b200/dichavit.py).
 *
 * The reference (chaudatascience/diverse_channel_vit) has no FFI: its hot path is
 * the Python nn.Module models/dichavit.py + models/vit.py + models/loss_fn.py that
 * dispatches to ATen.  Every entry point below therefore replaces an ATen call
 * site of that module; the reference file:line each one stands in for is cited.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the
 *     name ends in _host.  The caller owns every buffer, including workspaces.
 *   - `stream` is a cudaStream_t passed as void*.  Functions only enqueue work:
 *     no allocation, no synchronisation, no global mutable state.
 *   - return 0 on success, a negative DCV_ERR_* otherwise; dcv_last_error()
 *     returns a thread-local message for the last failure.
 *   - bf16 buffers are `uint16_t`-sized elements (void* here), fp32 are float*.
 */
#ifndef DCVIT_H_
#define DCVIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCV_OK 0
#define DCV_ERR_INVALID (-1)     /* bad argument (null pointer, empty shape, bad enum) */
#define DCV_ERR_UNSUPPORTED (-2) /* shape/alignment the kernels do not implement */
#define DCV_ERR_CUDA (-3)        /* a CUDA runtime/driver call failed */

const char* dcv_last_error(void);
int dcv_version(void);
/* number of kernels launched by this library since load (bench.py "gpu_launches") */
long long dcv_launch_count(void);

/* ---- epilogues of dcv_gemm_nt ---- */
#define DCV_EPI_BIAS 0       /* out(bf16) = A*B^T (+ bias)                                  */
#define DCV_EPI_BIAS_GELU 1  /* out(bf16) = h = A*B^T + bias ; out2(bf16) = gelu_erf(h)      */
#define DCV_EPI_BIAS_RESID 2 /* out(f32)  = resid(f32) + A*B^T + bias  (out may alias resid) */
#define DCV_EPI_DGELU 3      /* out(bf16) = (A*B^T) * gelu_erf'(aux(bf16))                  */
#define DCV_EPI_F32 4        /* out(f32)  = A*B^T (+ bias)                                  */

/* C[M,N] = A[M,K] * B[N,K]^T with a fused epilogue; A, B bf16 row-major.
 * Replaces nn.Linear forward (models/vit.py:116 qkv, :118 proj, :71 fc1 + :65 GELU,
 * :73 fc2, residual adds :397-398) and the autograd dgrad of the same layers. */
int dcv_gemm_nt(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue,
                const float* bias, void* out, void* out2, const float* resid, const void* aux, int ldo,
                void* stream);

/* C[M,N] = A[M,K] * B[K,N]; A bf16 [M,K], B bf16 [K,N] row-major (consumed MN-major, no transposed
 * copy).  epilogue in {DCV_EPI_BIAS (no bias: plain bf16 store), DCV_EPI_DGELU, DCV_EPI_F32}.
 * Replaces the autograd input gradient of nn.Linear (dX = dY W) and, with DCV_EPI_DGELU, the
 * GELU backward of models/vit.py:77-78. */
int dcv_gemm_nn(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue, void* out,
                const void* aux, int ldo, void* stream);

/* C[Nout,Kout] (+)= A[M,Nout]^T * B[M,Kout]; A, B bf16 row-major, C fp32.
 * accumulate=1: split-K atomic accumulation into C (caller zero-fills or holds a
 * running gradient); accumulate=0: plain store, single split.  splits<=0: auto.
 * Replaces the autograd weight gradient of nn.Linear (dW = dY^T X). */
int dcv_gemm_tn(const void* A, int lda, const void* B, int ldb, int M, int Nout, int Kout, float* C, int ldc,
                int accumulate, int splits, void* stream);

/* Multi-head self-attention forward, head_dim 64: o = softmax(q k^T * scale) v.
 * qkv bf16 [B, L, 3*H*64] (columns q|k|v, head h at h*64), o bf16 [B, L, H*64],
 * lse2 fp32 [B, H, Lp] (Lp = L rounded up to 128) = log2-domain log-sum-exp of the scaled scores (saved for backward).
 * Replaces models/vit.py:123-141 (q @ k^T * scale, softmax, attn @ v, transpose/reshape). */
int dcv_attn_fwd(const void* qkv, void* o, float* lse2, int B, int L, int H, float scale, void* stream);

/* Attention backward: dqkv bf16 [B, L, 3*H*64] from (qkv, o, dO, lse2).
 * Workspaces (caller-owned): delta fp32 [B, H, Lp]; dq_acc fp32 [B, H, L, 64] (zeroed inside).
 * Replaces the autograd backward of models/vit.py:126-141. */
int dcv_attn_bwd(const void* qkv, const void* o, const void* dO, const float* lse2, float* delta, float* dq_acc,
                 void* dqkv, int B, int L, int H, float scale, void* stream);

/* debug: override the MN-major shared-memory descriptor strides of dcv_gemm_tn (0 = default) */
void dcv_debug_set_tn_desc(int lbo_bytes, int sbo_bytes);

#ifdef __cplusplus
}
#endif
#endif /* DCVIT_H_ */
