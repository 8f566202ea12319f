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

// ---- built-in kernel profiler (api.cu): CUDA-event pairs around every launch of a tagged kernel
// class, recorded on the launching stream; off by default (one predictable branch per launch).
enum ProfTag {
  PT_GEMM_NT = 0,    // forward Linear (bias / gelu / residual epilogues)
  PT_GEMM_NN,        // dgrad
  PT_GEMM_TN,        // wgrad (split-K)
  PT_ATTN_FWD,
  PT_ATTN_BWD_PREP,  // delta = rowsum(dO o O) + dq accumulator clear
  PT_ATTN_BWD,
  PT_ATTN_BWD_FIN,   // fp32 dq accumulator -> bf16
  PT_LN_FWD,
  PT_LN_BWD,
  PT_COLSUM,
  PT_CAST,
  PT_IM2COL,
  PT_EMBED_GEMM,
  PT_EMBED_MISC,     // addend / CLS rows / extra-loss scalar
  PT_TDL,
  PT_CDL,
  PT_EMBED_BWD,      // dY assembly, batch sums, cls/pos/channel-token gradients
  PT_SMALL,          // fp32 SIMT GEMM (pos resample, head), CLS LayerNorm
  PT_COUNT
};
struct ProfScope {
  ProfScope(int tag, cudaStream_t st);
  ~ProfScope();
  int slot;
  cudaStream_t st;
};

// 2-D row-major tensor of bf16 (is_f32 = false) or fp32 elements, 128-byte-swizzled boxes
int make_tmap_2d(CUtensorMap* m, const void* ptr, bool is_f32, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer);
// 3-D fp32 tensor (used for the TMA reduce-add of the attention dQ accumulator)
int make_tmap_f32_3d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                     uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2);
// 5-D fp32 map without swizzle (image strips of the fused patch embedding); not cached
int make_tmap_f32_5d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t d4,
                     uint64_t s1, uint64_t s2, uint64_t s3, uint64_t s4, uint32_t b0, uint32_t b1, uint32_t b2,
                     uint32_t b3, uint32_t b4);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per (kernel, device): applied once for each pair, thread-safe
int ensure_smem_attr(const void* kernel, int bytes);
#define DCV_TRY_SMEM_ATTR(kern, bytes)                                            \
  do {                                                                            \
    int _rc = ::dcv::ensure_smem_attr(reinterpret_cast<const void*>(kern), bytes); \
    if (_rc != 0) return _rc;                                                     \
  } while (0)

#define DCV_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::dcv::set_error(DCV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                                  \
  } while (0)

// ---- gemm.cu ----
int gemm_nt(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epi, const float* bias,
            void* out, void* out2, const float* resid, const void* aux, int ldo, bool b_mn, cudaStream_t st,
            int map_T = 0, int map_L = 0, const float* addend = nullptr, int ldin = 0, float* delta = nullptr,
            int seq_L = 0);
int gemm_tn(const void* A, int lda, const void* B, int ldb, int M, int Nout, int Kout, float* C, int ldc,
            int accumulate, int splits, cudaStream_t st);
void debug_set_tn_desc(int lbo, int sbo);
void debug_set_nt_cluster(int cm);

// ---- attention.cu ----
// q_tiles > 0: only the first q_tiles 128-query tiles are computed (last block: only the CLS query is consumed)
int attn_fwd(const void* qkv, void* o, float* lse2, int B, int L, int H, float scale, cudaStream_t st,
             int q_tiles = 0);
// cls_only: dO is compact [B, D] (gradient of the CLS rows of o, every other row being zero); only query tile 0 is
// visited, dK / dV / dQ are still produced for all rows
// dbias_qkv: optional [3D] fp32, the column sums of dqkv (= qkv bias gradient) are ADDED to it
// dq_cleared: the caller has already zeroed dq_acc (the block backward does it on its side stream)
// delta_ready: delta[b,h,q] = sum_d dO*O was already produced (epilogue of the projection dgrad GEMM, EPI_DELTA; pad
// rows [L, Lp) zero); otherwise a prep kernel computes it here
int attn_bwd(const void* qkv, const void* o, const void* dO, const float* lse2, float* delta, float* dq_acc,
             void* dqkv, int B, int L, int H, float scale, cudaStream_t st, bool cls_only = false,
             bool delta_ready = false, float* dbias_qkv = nullptr, bool dq_cleared = false);

int debug_attn_timeline(long long* buf);
void debug_set_attn_fwd_mode(int m);
void debug_set_attn_bwd_mode(int m);

// ---- programmatic dependent launch ----
// Opt-in: DCV_PDL=1 in the environment (read once) or dcv_debug_set_pdl(1); otherwise plain stream ordering.
bool pdl_enabled();
void debug_set_pdl(int on);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- side stream (api.cu) ----
// Work of one launcher call that does not sit on its critical path (the weight-gradient GEMMs and the accumulator
// clears of the block backward) runs on a per-device side stream, forked from / joined into the caller's stream with
// events: under stream capture the fork / join become edges of the CUDA graph (a parallel branch), eagerly they are
// ordinary cross-stream dependencies.  side_branch() returns nullptr -- the caller then keeps everything on its own
// stream -- when the overlap is switched off (DCV_BWD_OVERLAP=0 / dcv_debug_set_bwd_overlap(0)), while the built-in
// profiler is recording (its per-class event pairs assume one kernel at a time), and when the stream / events of this
// device would have to be created while `main` is being captured.
struct SideBranch {
  cudaStream_t s;
  cudaEvent_t fork;
  cudaEvent_t join[4];
};
SideBranch* side_branch(cudaStream_t main, int rows);
void debug_set_bwd_overlap(int on);
// side waits for everything enqueued on `main` so far
inline cudaError_t side_fork(SideBranch* sb, cudaStream_t main) {
  cudaError_t e = cudaEventRecord(sb->fork, main);
  return e != cudaSuccess ? e : cudaStreamWaitEvent(sb->s, sb->fork, 0);
}
// `main` waits for everything enqueued on the side stream before join mark i was set
inline cudaError_t side_mark(SideBranch* sb, int i) { return cudaEventRecord(sb->join[i], sb->s); }
inline cudaError_t side_join(SideBranch* sb, int i, cudaStream_t main) { return cudaStreamWaitEvent(main, sb->join[i], 0); }

// ---- rowops.cu ----
int ln_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M, int D,
           float eps, cudaStream_t st);
int ln_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* gamma, float* dres,
           void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, int M, int D, cudaStream_t st);
int colsum_bf16(const void* a, float* out, int M, int N, int lda, cudaStream_t st);
int colsum_f32(const float* a, float* out, int M, int N, int lda, cudaStream_t st);
int cast_f32_bf16(const float* src, void* dst, long long n, cudaStream_t st);
int adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
               float beta2, float eps, float wd, int step, const float* clip, cudaStream_t st);
int sumsq_f32(const float* g, long long n, float* out, cudaStream_t st);
int adamw_step_dev(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float beta1, float beta2,
                   float eps, const dcv_optim_state* state, const float* clip, cudaStream_t st);
int optim_sched_step(dcv_optim_state* state, const dcv_sched& cfg, cudaStream_t st);

// ---- embed.cu ----
int im2col_gather(const void* x, int x_is_u8, const float* pix_mean, const float* pix_inv_std, const int* idx,
                  void* patches, int B, int C, int Cs, int H, int W, int P, cudaStream_t st);
int split_weight(const float* w, void* ws, int D, int K, cudaStream_t st);
int embed_addend(const float* bias, const float* chan_embed, const int* gid, const float* pos_patch, const float* cls,
                 const float* pos0, float* addend, float* tokens, int B, int Cs, int N, int D, cudaStream_t st);
int tdl_fwd(const float* tokens, const float* addend, const float* bias, float* S, float* Q, float* rnorm,
            float* S_all, float* loss_b, float* coef_pos, float* coef_neg, float* tdl_out, int B, int Cs, int N, int D,
            float gamma_s, float gamma_d, int reverse_pos_pairs, int use_square, cudaStream_t st);
int embed_bwd_dy(const float* G, const float* tokens, const float* addend, const float* bias, const float* rnorm,
                 const float* S, const float* S_all, const float* coef_pos, const float* coef_neg,
                 const float* d_extra, float lambda_tdl, void* dY, int B, int Cs, int N, int D, cudaStream_t st);
int embed_param_grads(const float* G, float* R, const int* gid, float* d_cls, float* d_pos0, float* d_chan_embed,
                      float* dpos_patch, int accumulate_pos, int B, int Cs, int N, int D, cudaStream_t st);
int cdl_fwd(const float* chan_embed, const float* proxies, const int* gid, float scale, float* loss, float* dE,
            float* dP, int Cs, int D, cudaStream_t st);
int cdl_bwd(const float* dE, const float* dP, const int* gid, const float* d_extra, float lambda_cdl,
            float* g_chan_embed, float* g_proxies, int Cs, int D, cudaStream_t st);
int extra_loss(const float* tdl, const float* cdl, float lt, float lc, float* extra, cudaStream_t st);
int sgemm_small(const float* A, int lda, int transA, const float* Bm, int ldb, int transB, float* C, int ldc,
                const float* bias, int accumulate, int M, int N, int K, cudaStream_t st);
int cls_ln_fwd(const float* x, long long row_stride, const float* gamma, const float* beta, float* feat, float* mean,
               float* rstd, int B, int D, float eps, cudaStream_t st);
int cls_ln_bwd(const float* dfeat, const float* x, long long row_stride, const float* mean, const float* rstd,
               const float* gamma, float* dres, void* dres_bf16, float* dgamma, float* dbeta, float* dxsum, int B,
               int D, cudaStream_t st);

// ---- embed_fused.cu ----
bool embed_fused_ok(const dcv_embed_dims& d, int x_is_u8);
void debug_set_embed_fused(int on);
int debug_embed_timeline(long long* buf);
int embed_fused_fwd(const dcv_embed_dims& d, const void* x, const int* idx, const void* wsplit, const float* bias,
                    const float* addend, float* tokens, void* patches, float* S, float* Q, float* rnorm, int tdl_on,
                    cudaStream_t st);
// the two small kernels that turn S / Q into the TDL scalar and its backward coefficients (second half of tdl_fwd)
int tdl_finish(float* S, float* Q, float* S_all, float* loss_b, float* coef_pos, float* coef_neg, float* tdl_out, int B,
               int Cs, int N, int D, float gamma_s, float gamma_d, int reverse_pos_pairs, int use_square, cudaStream_t st);

// ---- model.cu (stage orchestration behind dcv_block_* / dcv_embed_* / dcv_head_*) ----
int block_fwd(const dcv_dims& d, const dcv_block_params& p, const dcv_block_acts& a, cudaStream_t st);
int block_bwd(const dcv_dims& d, const dcv_block_params& p, const dcv_block_acts& a, const dcv_block_grads& g,
              const dcv_block_ws& ws, float* dres, void* dres_bf16, float* dbias_prev, cudaStream_t st);
int block_fwd_cls(const dcv_dims& d, const dcv_block_params& p, const dcv_block_acts& a, cudaStream_t st);
int block_bwd_cls(const dcv_dims& d, const dcv_block_params& p, const dcv_block_acts& a, const dcv_block_grads& g,
                  const dcv_block_ws& ws, float* dres_c, void* dres_c_bf16, float* dres, void* dres_bf16,
                  float* dbias_prev, cudaStream_t st);
int embed_fwd(const dcv_embed_dims& d, const dcv_embed_cfg& cfg, const dcv_embed_params& p, const void* x,
              const int* idx, const int* gid, const dcv_embed_acts& a, cudaStream_t st);
int embed_bwd(const dcv_embed_dims& d, const dcv_embed_cfg& cfg, const dcv_embed_params& p, const int* gid,
              const dcv_embed_acts& a, const dcv_embed_grads& g, const dcv_embed_ws& ws, const float* G,
              const float* d_extra, cudaStream_t st);
int head_fwd(const float* x_last, int B, int L, int D, const float* norm_w, const float* norm_b, float* feat,
             float* mean, float* rstd, const float* head_w, const float* head_b, float* logits, int num_classes,
             cudaStream_t st);
int head_bwd(const float* d_out, const float* x_last, int B, int L, int D, const float* norm_w, const float* feat,
             const float* mean, const float* rstd, const float* head_w, int num_classes, float* dfeat_ws, float* dres,
             void* dres_bf16, float* g_norm_w, float* g_norm_b, float* g_head_w, float* g_head_b, float* dbias_last,
             cudaStream_t st);

}  // namespace dcv
