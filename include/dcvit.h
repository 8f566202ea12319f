/* dcvit.h -- C ABI of libdcvit.so, the sm_100a kernel library behind the drop-in
 * DiChaViT module (diverse_channel_vit_b200/dichavit.py).
 *
 * The reference (chaudatascience/diverse_channel_vit) has no FFI: its hot path is
 * the Python nn.Module models/dichavit.py + models/vit.py + models/loss_fn.py that
 * dispatches to ATen.  Every entry point below therefore replaces an ATen call
 * site (or a group of them) of that module; the reference file:line each one stands
 * in for is cited.  Paths are relative to the reference repository root.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers.  The caller
 *     owns every buffer, including activations kept for backward and workspaces.
 *   - `stream` is a cudaStream_t passed as void*.  Functions only enqueue work: no device
 *     allocation, no synchronisation.  Process-wide state is limited to caches and switches that
 *     do not change results: per-(kernel, device) shared-memory opt-ins, a per-thread cache of
 *     encoded TMA descriptors, the launch counter / profiler, the dcv_debug_* A/B switches, and
 *     one side stream + a few events per device (dcv_block_bwd / dcv_embed_bwd / dcv_head_bwd fork
 *     work that is off their critical path onto it and join it back into `stream` before they
 *     return, so from the caller's point of view everything is still ordered on `stream`; under
 *     stream capture the fork / join become graph edges; one host thread per device).
 *   - return 0 on success, a negative DCV_ERR_* otherwise; dcv_last_error()
 *     returns a thread-local message for the last failure.
 *   - bf16 buffers are void*, fp32 are float*.  "[r, c]" is row-major.
 *   - token tensors are [B, L, D] with L = 1 + C'*N: row 0 CLS, row 1 + c*N + p =
 *     patch p of the c-th sampled channel (reference dichavit.py:414-415, :561-562).
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

/* =============================== primitive operators =============================== */

/* ---- epilogues of dcv_gemm_nt ---- */
#define DCV_EPI_BIAS 0       /* out(bf16) = A*B^T (+ bias)                                  */
#define DCV_EPI_BIAS_GELU 1  /* out(bf16) = h = A*B^T + bias ; out2(bf16) = gelu_erf(h)      */
#define DCV_EPI_BIAS_RESID 2 /* out(f32)  = resid(f32) + A*B^T + bias  (out may alias resid) */
#define DCV_EPI_DGELU 3      /* out(bf16) = (A*B^T) * gelu_erf'(aux(bf16))                  */
#define DCV_EPI_F32 4        /* out(f32)  = A*B^T (+ bias)                                  */

/* C[M,N] = A[M,K] * B[N,K]^T with a fused epilogue; A, B bf16 row-major.
 * Replaces nn.Linear forward (models/vit.py:116 qkv, :118 proj, :71 fc1 + :65 GELU,
 * :73 fc2, residual adds :397-398). */
int dcv_gemm_nt(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue,
                const float* bias, void* out, void* out2, const float* resid, const void* aux, int ldo,
                void* stream);

/* C[M,N] = A[M,K] * B[K,N]; A bf16 [M,K], B bf16 [K,N] row-major (consumed MN-major, no transposed
 * copy).  epilogue in {DCV_EPI_BIAS (no bias: plain bf16 store), DCV_EPI_DGELU, DCV_EPI_F32}.
 * Replaces the autograd input gradient of nn.Linear (dX = dY W) and, with DCV_EPI_DGELU, the
 * GELU backward of models/vit.py:77-78. */
int dcv_gemm_nn(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue, void* out,
                const void* aux, int ldo, void* stream);

/* Projection input gradient fused with the softmax-backward row term: dO[M,N] = dY[M,K] * W[K,N] (bf16 out) and
 * delta[b, h, q] = sum over head h's 64 columns of dO * O, rows m = b*L + q, delta fp32 [B, N/64, Lp]
 * (Lp = L rounded up to 128; the caller zero-fills the pad rows).  Replaces the autograd backward of
 * models/vit.py:140-141 (proj Linear dgrad) plus the rowsum(dO o O) pass of the attention backward. */
int dcv_gemm_nn_delta(const void* dY, int lda, const void* W, int ldb, int M, int N, int K, void* dO, const void* O,
                      float* delta, int L, void* stream);

/* C[Nout,Kout] (+)= A[M,Nout]^T * B[M,Kout]; A, B bf16 row-major, C fp32.
 * accumulate=1: split-K atomic accumulation into C (caller zero-fills or holds a
 * running gradient); accumulate=0: plain store, single split.  splits<=0: auto.
 * Replaces the autograd weight gradient of nn.Linear (dW = dY^T X). */
int dcv_gemm_tn(const void* A, int lda, const void* B, int ldb, int M, int Nout, int Kout, float* C, int ldc,
                int accumulate, int splits, void* stream);

/* Multi-head self-attention forward, head_dim 64: o = softmax(q k^T * scale) v.
 * qkv bf16 [B, L, 3*H*64] (columns q|k|v, head h at h*64), o bf16 [B, L, H*64],
 * lse2 fp32 [B, H, Lp] (Lp = L rounded up to 128) = log2-domain log-sum-exp of the scaled
 * scores (saved for backward).
 * Replaces models/vit.py:123-141 (q @ k^T * scale, softmax, attn @ v, transpose/reshape). */
int dcv_attn_fwd(const void* qkv, void* o, float* lse2, int B, int L, int H, float scale, void* stream);

/* Attention backward: dqkv bf16 [B, L, 3*H*64] from (qkv, o, dO, lse2).
 * Workspaces (caller-owned): delta fp32 [B, H, Lp]; dq_acc fp32 [B, H, L, 64] (zeroed inside).
 * Replaces the autograd backward of models/vit.py:126-141. */
int dcv_attn_bwd(const void* qkv, const void* o, const void* dO, const float* lse2, float* delta, float* dq_acc,
                 void* dqkv, int B, int L, int H, float scale, void* stream);
/* Same, with the two fusions the block backward uses: delta_ready != 0 -> delta[b,h,q] was already produced by
 * dcv_gemm_nn_delta (no prep pass); dbias_qkv != NULL (fp32 [3*H*64]) -> the column sums of dqkv,
 * i.e. the qkv bias gradient (autograd of models/vit.py:121), are ADDED to it by the kernels' epilogues. */
int dcv_attn_bwd_fused(const void* qkv, const void* o, const void* dO, const float* lse2, float* delta, float* dq_acc,
                       void* dqkv, float* dbias_qkv, int delta_ready, int B, int L, int H, float scale, void* stream);


/* LayerNorm forward over rows (nn.LayerNorm, biased variance): x fp32 [M,D] -> y bf16 [M,D],
 * mean/rstd fp32 [M].  Replaces models/vit.py:384,398 (norm1, norm2; eps 1e-6 dichavit.py:724). */
int dcv_ln_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M,
               int D, float eps, void* stream);

/* LayerNorm backward fused with the residual-stream gradient add:
 *   dres(fp32, in/out) += LN'(dy);  dx_bf16 = bf16(dres);  dgamma += ..; dbeta += ..;
 *   dxsum (nullable) += column sums of the updated dres (bias gradient of the Linear that
 *   produced this residual branch's input).  Replaces autograd of the same lines. */
int dcv_ln_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
               float* dres, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, int M, int D, void* stream);

/* out[n] += sum_m a[m,n], a bf16 [M,N] (bias gradients of qkv / fc1). */
int dcv_colsum_bf16(const void* a, float* out, int M, int N, int lda, void* stream);

/* dst(bf16)[i] = src(fp32)[i]: the per-step bf16 copy of the flat parameter buffer. */
int dcv_cast_f32_bf16(const float* src, void* dst, long long n, void* stream);

/* Fused AdamW over flat fp32 buffers of n elements (n % 4 == 0): timm/torch AdamW semantics of reference
 * optimizers.py:20-21 (decoupled weight decay, bias correction with `step` counted from 1).  p_bf16 (nullable):
 * also refreshes the bf16 operand copy.  clip (nullable): device float[2] = {sum of squared gradients, max_norm};
 * gradients are scaled by min(1, max_norm / (sqrt(sumsq) + 1e-6)) as torch clip_grad_norm_ (trainer.py:1003-1004). */
int dcv_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, const float* clip, void* stream);
/* ---- device-resident optimiser scalars (CUDA-graph friendly: a captured step replays with the values of the
 * current update).  dcv_optim_sched_step advances state->num_updates by one and evaluates, ON THE DEVICE,
 *   - the learning rate of timm's CosineLRScheduler._get_lr (reference lr_schedulers.py:6-9; parameters of
 *     configs/scheduler/cosine.yaml; cycle_mul == 1, no noise) -- t counts updates (updates_per_epoch == 0: the
 *     lr set by step_update(num_updates) after the previous optimizer.step(), trainer.py:1009-1010) or 1-based
 *     epochs (updates_per_epoch > 0: scheduler.step(epoch), trainer.py:344-348); t_initial <= 0: constant base_lr;
 *   - the weight decay of utils.cosine_scheduler (utils.py:563-574) as indexed by trainer.py:1011-1019
 *     (wd_total = epochs * updates_per_epoch entries; wd_total == 0: constant wd_base);
 *   - Adam's bias corrections for update number state->num_updates.
 * dcv_adamw_step_dev is dcv_adamw_step reading lr / weight decay / bias corrections from that state. ---- */
typedef struct dcv_optim_state { /* device memory, 32 bytes; zero-filled = no update applied yet */
  int num_updates;
  float lr, wd, bc1, bc2_sqrt;
  float reserved[3];
} dcv_optim_state;

typedef struct dcv_sched {
  float base_lr, lr_min, warmup_lr_init;
  int t_initial, warmup_t, warmup_prefix, cycle_limit;
  float cycle_decay, k_decay;
  int updates_per_epoch;
  float wd_base, wd_end;
  int wd_total;
  float beta1, beta2;
} dcv_sched;

int dcv_optim_sched_step(dcv_optim_state* state, const dcv_sched* cfg, void* stream);
int dcv_adamw_step_dev(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float beta1, float beta2,
                       float eps, const dcv_optim_state* state, const float* clip, void* stream);

/* out[0] += sum_i g[i]^2 (n % 4 == 0; caller zeroes out) */
int dcv_sumsq_f32(const float* g, long long n, float* out, void* stream);

/* C[M,N] = (accumulate ? C : 0) + op(A) op(B) (+ bias[N]); tiny fp32 SIMT GEMM used for the
 * bicubic positional-embedding resample (dichavit.py:518-552, a fixed linear map) and the
 * classifier head (dichavit.py:801,855).  transA: A stored [K,M]; transB: B stored [N,K]. */
int dcv_sgemm_small(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C, int ldc,
                    const float* bias, int accumulate, int M, int N, int K, void* stream);

/* =============================== fused stages =============================== */

/* ---- transformer block (models/vit.py:346-399 Block.forward and its autograd) ---- */
typedef struct dcv_dims {
  int B, L, D, H, F; /* images, tokens per image, embed dim, heads (D = 64 H), MLP hidden */
} dcv_dims;

typedef struct dcv_block_params { /* fp32 vectors; bf16 [out, in] weight copies */
  const float *ln1_w, *ln1_b, *qkv_b, *proj_b, *ln2_w, *ln2_b, *fc1_b, *fc2_b;
  const void *qkv_w, *proj_w, *fc1_w, *fc2_w;
} dcv_block_params;

typedef struct dcv_block_grads { /* fp32, ACCUMULATED into (caller zero-fills once per step) */
  float *ln1_w, *ln1_b, *qkv_w, *qkv_b, *proj_w, *proj_b, *ln2_w, *ln2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} dcv_block_grads;

typedef struct dcv_block_acts { /* activations written by forward, read by backward; M = B*L rows */
  const float* x_in; /* fp32 [M, D] block input (owned by the previous stage)       */
  void* u;           /* bf16 [M, D]  LN1 output                                      */
  float *mean1, *rstd1;
  void* qkv;   /* bf16 [M, 3D]                                                        */
  void* o;     /* bf16 [M, D]  attention output (head-major columns)                  */
  float* lse2; /* fp32 [B, H, Lp]                                                     */
  float* x_mid; /* fp32 [M, D]  after the attention residual                          */
  void* v;      /* bf16 [M, D]  LN2 output                                            */
  float *mean2, *rstd2;
  void* h;      /* bf16 [M, F]  fc1 pre-activation                                    */
  void* g;      /* bf16 [M, F]  gelu(h)                                               */
  float* x_out; /* fp32 [M, D]  block output                                          */
} dcv_block_acts;

typedef struct dcv_block_ws { /* backward scratch, reusable across blocks */
  void* dh;      /* bf16 [M, F]  */
  void* dv;      /* bf16 [M, D]  (also reused for du) */
  void* d_o;     /* bf16 [M, D]  */
  void* dqkv;    /* bf16 [M, 3D] */
  float* delta;  /* fp32 [B, H, Lp] */
  float* dq_acc; /* fp32 [B, H, L, 64] */
} dcv_block_ws;

int dcv_block_fwd(const dcv_dims* dims, const dcv_block_params* p, const dcv_block_acts* a, void* stream);

/* dres fp32 [M,D] / dres_bf16 [M,D]: gradient w.r.t. the block output on entry, w.r.t. the block
 * input on exit.  dbias_prev (nullable): receives += column sums of the exit gradient, i.e. the
 * bias gradient of the previous block's fc2.  This block's fc2 bias gradient is produced by the
 * NEXT stage in backward order (the later block's dbias_prev / dcv_head_bwd). */
int dcv_block_bwd(const dcv_dims* dims, const dcv_block_params* p, const dcv_block_acts* a,
                  const dcv_block_grads* g, const dcv_block_ws* ws, float* dres, void* dres_bf16,
                  float* dbias_prev, void* stream);

/* Last transformer block, exploiting that only the CLS row of its output is consumed (dichavit.py:651-652:
 * `x = self.norm(x); return x[:, 0]`): LN1, the qkv projection and K/V stay full size, attention runs for
 * the first 128-query tile only, and the output projection, LN2 and the MLP run on one row per image.  Results
 * on the consumed rows are identical to dcv_block_fwd/bwd.  In `a`, the tensors after the attention are
 * COMPACT: x_mid, v, mean2, rstd2, h, g, x_out hold B rows (not B*L); u, mean1, rstd1, qkv, o, lse2 are full
 * size (o / lse2 are written for query rows 0..127 of every image only). */
int dcv_block_fwd_cls(const dcv_dims* dims, const dcv_block_params* p, const dcv_block_acts* a, void* stream);

/* dres_c fp32 [B, D] (+ bf16 copy dres_c_bf16): gradient w.r.t. the CLS rows of the block output (all other rows
 * are zero), clobbered.  On exit dres [M, D] / dres_bf16 [M, D] hold the gradient w.r.t. the whole block input.
 * ws: dh and d_o need B rows only; dv [M, D], dqkv [M, 3D], delta, dq_acc as in dcv_block_bwd. */
int dcv_block_bwd_cls(const dcv_dims* dims, const dcv_block_params* p, const dcv_block_acts* a,
                      const dcv_block_grads* g, const dcv_block_ws* ws, float* dres_c, void* dres_c_bf16,
                      float* dres, void* dres_bf16, float* dbias_prev, void* stream);

/* ---- channel-adaptive patch embedding + DCS gather + CLS/pos + TDL + CDL ----
 * models/dichavit.py:110-417 (PatchEmbedPerChannel.forward: x[:, idx] gather :210, Conv3d proj :377,
 * TDL :378-389, CDL :399-402, extra loss :406-408, + channel_embed :409-411), :554-565 (CLS, pos),
 * :518-552 (bicubic pos resample), models/loss_fn.py:7-59. */
typedef struct dcv_embed_dims {
  int B, C, Cs, H, W, P, D; /* images, channels in x, sampled channels C', image H x W, patch, embed dim */
} dcv_embed_dims;

typedef struct dcv_embed_cfg {
  float lambda_tdl, lambda_cdl; /* ortho_loss_v1_lambda, proxy_loss_lambda (0 = loss off) */
  float gamma_s, gamma_d;
  float cdl_scale; /* sqrt(1 / temperature), dichavit.py:60 */
  int reverse_pos_pairs, use_square;
  int x_is_u8; /* 1: x holds raw uint8 pixels, standardised on the device with pix_mean / pix_inv_std */
} dcv_embed_cfg;

typedef struct dcv_embed_params {
  const float* proj_w;     /* fp32 [D, P*P] (Conv3d weight reshaped)                       */
  const float* proj_b;     /* [D]                                                          */
  const float* chan_embed; /* [C_total, D] channel_embed.weight                            */
  const float* proxies;    /* [C_total, D] channel_emb_proxies (NULL if CDL off)           */
  const float* cls;        /* [D]                                                          */
  const float* pos;        /* [1 + N, D] pos_embed                                         */
  const float* pos_map;    /* [N, N] bicubic resample matrix, NULL = use pos[1:] unchanged */
  const float* pix_mean;   /* [C] per input-channel mean of the loader's standardisation (x_is_u8 only; NULL = none) */
  const float* pix_inv_std; /* [C] 1 / std                                                                      */
} dcv_embed_params;

typedef struct dcv_embed_grads { /* fp32, accumulated */
  float *proj_w, *proj_b, *chan_embed, *proxies, *cls, *pos;
} dcv_embed_grads;

typedef struct dcv_embed_acts {
  void* patches;    /* bf16 [B*C'*N, 3*P*P] gathered im2col rows, split precision [hi | lo | hi]  */
  void* wsplit;     /* bf16 [D, 3*P*P] split conv weight [Whi | Whi | Wlo]                        */
  float* pos_patch; /* fp32 [N, D]   */
  float* addend;    /* fp32 [C'*N, D] bias + channel token + positional embedding            */
  float* tokens;    /* fp32 [B, L, D] output                                                 */
  float *S, *Q, *rnorm, *S_all, *loss_b, *coef_pos, *coef_neg; /* TDL state: [B,C',D] [B,C'] [B,T] [B,D] [B] [B] [B] */
  float *cdl_dE, *cdl_dP; /* [C', D] unscaled CDL gradients                                 */
  float *tdl, *cdl, *extra; /* scalars: TDL, CDL, lambda_tdl*TDL + lambda_cdl*CDL            */
} dcv_embed_acts;

typedef struct dcv_embed_ws {
  void* dY;          /* bf16 [B*C'*N, D] */
  float* R;          /* fp32 [L, D]      */
  float* dpos_patch; /* fp32 [N, D]      */
} dcv_embed_ws;

/* x fp32 (or uint8 when cfg->x_is_u8: the loader's (x - mean) / std of datasets/dataset_utils.py:44 /
 * jump_cp_transforms.py:119-121 then runs on the device) [B, C, H, W]; idx int32 [C'] positions of the
 * sampled channels inside x (NULL = 0..C'-1); gid int32 [C'] their global channel ids (rows of chan_embed / proxies). */
int dcv_embed_fwd(const dcv_embed_dims* dims, const dcv_embed_cfg* cfg, const dcv_embed_params* p, const void* x,
                  const int* idx, const int* gid, const dcv_embed_acts* a, void* stream);

/* G fp32 [B, L, D]: gradient w.r.t. tokens; d_extra: device scalar, gradient w.r.t. the extra loss. */
int dcv_embed_bwd(const dcv_embed_dims* dims, const dcv_embed_cfg* cfg, const dcv_embed_params* p, const int* gid,
                  const dcv_embed_acts* a, const dcv_embed_grads* g, const dcv_embed_ws* ws, const float* G,
                  const float* d_extra, void* stream);

/* ---- final norm on the CLS row + classifier head (dichavit.py:651-652, :801, :855) ----
 * head_w NULL: out = feat (CHAMMI, Identity head); else logits = feat head_w^T + head_b. */
/* x_last rows are L*D apart (L = 1 for the compact output of dcv_block_fwd_cls). */
int dcv_head_fwd(const float* x_last, int B, int L, int D, const float* norm_w, const float* norm_b, float* feat,
                 float* mean, float* rstd, const float* head_w, const float* head_b, float* logits, int num_classes,
                 void* stream);

/* d_out fp32 [B, num_classes] (or [B, D] without head).  Zero-fills dres / dres_bf16 [B*L, D] and
 * writes their CLS rows; accumulates g_norm_w/b, g_head_w/b and (nullable) dbias_last = fc2 bias
 * gradient of the last block.  dfeat_ws fp32 [B, D] scratch (unused without head). */
int dcv_head_bwd(const float* d_out, const float* x_last, int B, int L, int D, const float* norm_w,
                 const float* feat, const float* mean, const float* rstd, const float* head_w, int num_classes,
                 float* dfeat_ws, float* dres, void* dres_bf16, float* g_norm_w, float* g_norm_b, float* g_head_w,
                 float* g_head_b, float* dbias_last, void* stream);

/* ---- built-in profiler: CUDA-event pairs around every kernel launch of each kernel class, on the
 * launching stream.  start() arms it; stop() synchronises the device (profiling only -- never on the
 * training path) and returns total milliseconds and scope counts per class. ---- */
int dcv_profile_num_tags(void);
const char* dcv_profile_tag_name(int tag);
int dcv_profile_start(void);
int dcv_profile_stop(double* ms_by_tag, long long* launches_by_tag, int ntags);

/* debug: GEMM cluster mode of dcv_gemm_nt / dcv_gemm_nn: 1 = single-CTA tiles (default), 2 = TMA multicast of the B
 * operand inside 2-CTA clusters, 3 = cta_group::2 pair MMA (256 x BN tiles) for K-major B */
void dcv_debug_set_nt_cluster(int cm);

/* debug: clock64() timeline of one CTA of the attention-backward kernel into buf (>= 3*1024 int64, device);
 * NULL switches it off */
int dcv_debug_attn_timeline(long long* buf);

/* debug / tuning: instruction-stream variants of the attention kernels (same results up to the documented exp2
 * polynomial error): 0 = scalar fp32 math, every exponential on the MUFU; 1 = packed-pair fp32 math (FFMA2 / FADD2,
 * 3-input max); 2, 3, 4 = 1 + 2 / 3 / 4 of every 8 exponential pairs evaluated by a polynomial on the FMA pipe.
 * A negative value leaves that kernel's mode unchanged. */
void dcv_debug_set_attn_mode(int fwd_mode, int bwd_mode);

/* debug / A-B timing: programmatic dependent launch of the GEMM, LayerNorm and attention kernels (off by default,
 * DCV_PDL=1 in the environment switches it on): each of them may become resident while its predecessor in the stream
 * drains and waits (griddepcontrol.wait) for the predecessor's results before touching global memory.  Measured
 * (tools/pdl_ab.py): faster only for the shortest sequences, slower from ~800 tokens up -- hence opt-in. */
void dcv_debug_set_pdl(int on);

/* debug / A-B timing: the block backward runs its weight-gradient GEMMs and accumulator clears on a side stream of the
 * library (a parallel branch of the graph when the call is captured), joined back before dcv_block_bwd returns to the
 * caller's stream; dcv_embed_bwd does the same with its cls / pos / channel-token gradient chain and dcv_head_bwd with
 * its two full-size clears.  0 = everything on the caller's stream, 1 = always, n > 1 = only for calls of at most n token rows,
 * < 0 = back to the default (DCV_BWD_OVERLAP in the environment overrides the default). */
void dcv_debug_set_bwd_overlap(int on);

/* debug / A-B timing of the patch embedding: 0 = the three-kernel path (gather + GEMM + TDL sums), 1 = the fused TMA-fed
 * kernel with one tile per CTA, 2 = the fused kernel as a persistent, cross-tile pipelined kernel, < 0 = back to the
 * default (DCV_EMBED_FUSED=0|1|2 in the environment overrides the default).  The fused kernels need P = 16, D = 384 and
 * fp32 input; other shapes always take the three-kernel path. */
void dcv_debug_set_embed_fused(int on);

/* debug: when buf != NULL (device memory, >= 128 int64), one CTA of the fused patch-embedding kernel records clock64()
 * stamps of its pipeline stages into it (tools/embed_timeline.py); NULL switches the recording off */
int dcv_debug_embed_timeline(long long* buf);

/* debug: override the MN-major shared-memory descriptor strides of dcv_gemm_tn (0 = default) */
void dcv_debug_set_tn_desc(int lbo_bytes, int sbo_bytes);

#ifdef __cplusplus
}
#endif
#endif /* DCVIT_H_ */
