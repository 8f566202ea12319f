// Stage orchestration: the launch sequences of one transformer block (forward / backward),
// of the channel-adaptive patch embedding with its two diversification losses, and of the
// CLS head.  Pure host code: every arithmetic step is one of the kernels in gemm.cu,
// attention.cu, rowops.cu, embed.cu.  Buffers are all caller-owned (include/dcvit.h).
#include "host.h"

namespace dcv {

enum { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RESID = 2, EPI_DGELU = 3, EPI_F32 = 4, EPI_EMBED = 5, EPI_DELTA = 6 };

#define DCV_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

static int check_dims(const dcv_dims& d) {
  if (d.B <= 0 || d.L <= 0 || d.D <= 0 || d.H <= 0 || d.F <= 0) return set_error(DCV_ERR_INVALID, "block: empty dims");
  if (d.D != d.H * 64) return set_error(DCV_ERR_UNSUPPORTED, "block: head_dim must be 64 (D=%d, H=%d)", d.D, d.H);
  if (d.D % 64 || d.F % 64) return set_error(DCV_ERR_UNSUPPORTED, "block: D and F must be multiples of 64");
  return 0;
}

// reference models/vit.py:383-399:  x += proj(attn(LN1 x));  x += fc2(gelu(fc1(LN2 x)))
int block_fwd(const dcv_dims& d, const dcv_block_params& p, const dcv_block_acts& a, cudaStream_t st) {
  DCV_TRY(check_dims(d));
  const int M = d.B * d.L, D = d.D, F = d.F;
  DCV_TRY(ln_fwd(a.x_in, p.ln1_w, p.ln1_b, a.u, a.mean1, a.rstd1, M, D, 1e-6f, st));
  DCV_TRY(gemm_nt(a.u, D, p.qkv_w, D, M, 3 * D, D, EPI_BIAS, p.qkv_b, a.qkv, nullptr, nullptr, nullptr, 3 * D, false, st));
  DCV_TRY(attn_fwd(a.qkv, a.o, a.lse2, d.B, d.L, d.H, 0.125f, st));
  DCV_TRY(gemm_nt(a.o, D, p.proj_w, D, M, D, D, EPI_BIAS_RESID, p.proj_b, a.x_mid, nullptr, a.x_in, nullptr, D, false, st));
  DCV_TRY(ln_fwd(a.x_mid, p.ln2_w, p.ln2_b, a.v, a.mean2, a.rstd2, M, D, 1e-6f, st));
  DCV_TRY(gemm_nt(a.v, D, p.fc1_w, D, M, F, D, EPI_BIAS_GELU, p.fc1_b, a.h, a.g, nullptr, nullptr, F, false, st));
  DCV_TRY(gemm_nt(a.g, F, p.fc2_w, F, M, D, F, EPI_BIAS_RESID, p.fc2_b, a.x_out, nullptr, a.x_mid, nullptr, D, false, st));
  return 0;
}

int block_bwd(const dcv_dims& d, const dcv_block_params& p, const dcv_block_acts& a, const dcv_block_grads& g,
              const dcv_block_ws& ws, float* dres, void* dres_bf16, float* dbias_prev, cudaStream_t st) {
  DCV_TRY(check_dims(d));
  const int M = d.B * d.L, D = d.D, F = d.F;
  // The chain dgrad -> LayerNorm backward -> attention backward -> dgrad is the critical path; the four weight
  // gradients and the two accumulator clears only have to be complete when their inputs are overwritten (dres_bf16 by
  // the LayerNorm backward passes) resp. when the block returns.  They go to the side stream (sb != nullptr), w = the
  // stream they are enqueued on.
  SideBranch* sb = side_branch(st, M);
  const cudaStream_t w = sb ? sb->s : st;
  const size_t delta_bytes = static_cast<size_t>(d.B) * d.H * ((d.L + 127) / 128 * 128) * sizeof(float);
  const size_t dq_bytes = static_cast<size_t>(d.B) * d.H * d.L * 64 * sizeof(float);
  if (sb) {
    DCV_CUDA(side_fork(sb, st));
    // dW2 += dres^T g                    [D,F]
    DCV_TRY(gemm_tn(dres_bf16, D, a.g, F, M, D, F, g.fc2_w, F, 1, 0, w));
    if (d.L % 128) DCV_CUDA(cudaMemsetAsync(ws.delta, 0, delta_bytes, w));
    DCV_CUDA(cudaMemsetAsync(ws.dq_acc, 0, dq_bytes, w));
    DCV_CUDA(side_mark(sb, 0));
  }
  // ---- MLP branch:  x_out = x_mid + fc2(gelu(fc1(LN2 x_mid))) ----
  // dh = (dres W2) o gelu'(h)            [M,F]
  // (the column sums of dh = fc1 bias gradient are accumulated by the same epilogue)
  DCV_TRY(gemm_nt(dres_bf16, D, p.fc2_w, F, M, F, D, EPI_DGELU, nullptr, ws.dh, nullptr, nullptr, a.h, F, true, st, 0, 0,
                  nullptr, 0, g.fc1_b));
  if (!sb) DCV_TRY(gemm_tn(dres_bf16, D, a.g, F, M, D, F, g.fc2_w, F, 1, 0, st));
  // dv = dh W1                           [M,D]
  if (sb) DCV_CUDA(side_fork(sb, st));  // dh is complete
  DCV_TRY(gemm_nt(ws.dh, F, p.fc1_w, D, M, D, F, EPI_BIAS, nullptr, ws.dv, nullptr, nullptr, nullptr, D, true, st));
  // dW1 += dh^T v
  DCV_TRY(gemm_tn(ws.dh, F, a.v, D, M, F, D, g.fc1_w, D, 1, 0, w));
  // dres += LN2'(dv); column sums of the result = d proj bias.  Overwrites dres_bf16: dW2 has to be through with it
  if (sb) DCV_CUDA(side_join(sb, 0, st));
  DCV_TRY(ln_bwd(ws.dv, a.x_mid, a.mean2, a.rstd2, p.ln2_w, dres, dres_bf16, g.ln2_w, g.ln2_b, g.proj_b, M, D, st));
  // ---- attention branch:  x_mid = x_in + proj(attn(qkv(LN1 x_in))) ----
  // dO = dres Wproj, with delta[b,h,q] = sum_d dO*O (the softmax-backward row term) from the same epilogue
  if (!sb && d.L % 128) DCV_CUDA(cudaMemsetAsync(ws.delta, 0, delta_bytes, st));
  if (sb) DCV_CUDA(side_fork(sb, st));  // the new dres_bf16 is complete
  DCV_TRY(gemm_nt(dres_bf16, D, p.proj_w, D, M, D, D, EPI_DELTA, nullptr, ws.d_o, nullptr, nullptr, a.o, D, true, st, 0, 0,
                  nullptr, 0, ws.delta, d.L));
  DCV_TRY(gemm_tn(dres_bf16, D, a.o, D, M, D, D, g.proj_w, D, 1, 0, w));
  if (sb) DCV_CUDA(side_mark(sb, 1));
  // the qkv bias gradient (column sums of dqkv) comes out of the attention-backward epilogues
  float* fused_db = g.qkv_b;
  DCV_TRY(attn_bwd(a.qkv, a.o, ws.d_o, a.lse2, ws.delta, ws.dq_acc, ws.dqkv, d.B, d.L, d.H, 0.125f, st, false, true,
                   fused_db, sb != nullptr));
  if (sb) DCV_CUDA(side_fork(sb, st));  // dqkv is complete
  DCV_TRY(gemm_nt(ws.dqkv, 3 * D, p.qkv_w, D, M, D, 3 * D, EPI_BIAS, nullptr, ws.dv, nullptr, nullptr, nullptr, D, true, st));
  DCV_TRY(gemm_tn(ws.dqkv, 3 * D, a.u, D, M, 3 * D, D, g.qkv_w, D, 1, 0, w));
  if (!fused_db) DCV_TRY(colsum_bf16(ws.dqkv, g.qkv_b, M, 3 * D, 3 * D, st));
  // overwrites dres_bf16: the projection weight gradient has to be through with it
  if (sb) DCV_CUDA(side_join(sb, 1, st));
  DCV_TRY(ln_bwd(ws.dv, a.x_in, a.mean1, a.rstd1, p.ln1_w, dres, dres_bf16, g.ln1_w, g.ln1_b, dbias_prev, M, D, st));
  if (sb) {  // everything of this call is ordered before whatever the caller enqueues next
    DCV_CUDA(side_mark(sb, 2));
    DCV_CUDA(side_join(sb, 2, st));
  }
  return 0;
}

// Last block: only the CLS row of the output is consumed (reference dichavit.py:651-652).  Everything after the
// attention runs on B gathered rows (row stride L*D in the full-size tensors), attention on the first query tile.
int block_fwd_cls(const dcv_dims& d, const dcv_block_params& p, const dcv_block_acts& a, cudaStream_t st) {
  DCV_TRY(check_dims(d));
  const int M = d.B * d.L, D = d.D, F = d.F, B = d.B, LD = d.L * d.D;
  DCV_TRY(ln_fwd(a.x_in, p.ln1_w, p.ln1_b, a.u, a.mean1, a.rstd1, M, D, 1e-6f, st));
  DCV_TRY(gemm_nt(a.u, D, p.qkv_w, D, M, 3 * D, D, EPI_BIAS, p.qkv_b, a.qkv, nullptr, nullptr, nullptr, 3 * D, false, st));
  DCV_TRY(attn_fwd(a.qkv, a.o, a.lse2, d.B, d.L, d.H, 0.125f, st, 1));
  // x_mid[b] = x_in[b, 0] + o[b, 0] Wproj^T + b : A and the residual are gathered with row stride L*D
  DCV_TRY(gemm_nt(a.o, LD, p.proj_w, D, B, D, D, EPI_BIAS_RESID, p.proj_b, a.x_mid, nullptr, a.x_in, nullptr, D, false, st,
                  0, 0, nullptr, LD));
  DCV_TRY(ln_fwd(a.x_mid, p.ln2_w, p.ln2_b, a.v, a.mean2, a.rstd2, B, D, 1e-6f, st));
  DCV_TRY(gemm_nt(a.v, D, p.fc1_w, D, B, F, D, EPI_BIAS_GELU, p.fc1_b, a.h, a.g, nullptr, nullptr, F, false, st));
  DCV_TRY(gemm_nt(a.g, F, p.fc2_w, F, B, D, F, EPI_BIAS_RESID, p.fc2_b, a.x_out, nullptr, a.x_mid, nullptr, D, false, st));
  return 0;
}

int block_bwd_cls(const dcv_dims& d, const dcv_block_params& p, const dcv_block_acts& a, const dcv_block_grads& g,
                  const dcv_block_ws& ws, float* dres_c, void* dres_c_bf16, float* dres, void* dres_bf16,
                  float* dbias_prev, cudaStream_t st) {
  DCV_TRY(check_dims(d));
  const int M = d.B * d.L, D = d.D, F = d.F, B = d.B, LD = d.L * d.D;
  // ---- MLP branch on the CLS rows ----
  DCV_TRY(gemm_nt(dres_c_bf16, D, p.fc2_w, F, B, F, D, EPI_DGELU, nullptr, ws.dh, nullptr, nullptr, a.h, F, true, st, 0, 0,
                  nullptr, 0, g.fc1_b));
  DCV_TRY(gemm_tn(dres_c_bf16, D, a.g, F, B, D, F, g.fc2_w, F, 1, 0, st));
  DCV_TRY(gemm_nt(ws.dh, F, p.fc1_w, D, B, D, F, EPI_BIAS, nullptr, ws.dv, nullptr, nullptr, nullptr, D, true, st));
  DCV_TRY(gemm_tn(ws.dh, F, a.v, D, B, F, D, g.fc1_w, D, 1, 0, st));
  DCV_TRY(ln_bwd(ws.dv, a.x_mid, a.mean2, a.rstd2, p.ln2_w, dres_c, dres_c_bf16, g.ln2_w, g.ln2_b, g.proj_b, B, D, st));
  // ---- attention branch: only the CLS rows of the projection input carry gradient ----
  DCV_TRY(gemm_nt(dres_c_bf16, D, p.proj_w, D, B, D, D, EPI_BIAS, nullptr, ws.d_o, nullptr, nullptr, nullptr, D, true, st));
  DCV_TRY(gemm_tn(dres_c_bf16, D, a.o, LD, B, D, D, g.proj_w, D, 1, 0, st));
  float* fused_db = g.qkv_b;
  DCV_TRY(attn_bwd(a.qkv, a.o, ws.d_o, a.lse2, ws.delta, ws.dq_acc, ws.dqkv, d.B, d.L, d.H, 0.125f, st, true, false,
                   fused_db));
  DCV_TRY(gemm_nt(ws.dqkv, 3 * D, p.qkv_w, D, M, D, 3 * D, EPI_BIAS, nullptr, ws.dv, nullptr, nullptr, nullptr, D, true, st));
  DCV_TRY(gemm_tn(ws.dqkv, 3 * D, a.u, D, M, 3 * D, D, g.qkv_w, D, 1, 0, st));
  if (!fused_db) DCV_TRY(colsum_bf16(ws.dqkv, g.qkv_b, M, 3 * D, 3 * D, st));
  // gradient w.r.t. the block input through the residual path: zero except the CLS rows
  DCV_CUDA(cudaMemsetAsync(dres, 0, static_cast<size_t>(M) * D * sizeof(float), st));
  DCV_CUDA(cudaMemcpy2DAsync(dres, static_cast<size_t>(LD) * sizeof(float), dres_c, static_cast<size_t>(D) * sizeof(float),
                             static_cast<size_t>(D) * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
  DCV_TRY(ln_bwd(ws.dv, a.x_in, a.mean1, a.rstd1, p.ln1_w, dres, dres_bf16, g.ln1_w, g.ln1_b, dbias_prev, M, D, st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// patch embedding
// ---------------------------------------------------------------------------------------------
static int check_embed(const dcv_embed_dims& d) {
  if (d.B <= 0 || d.C <= 0 || d.Cs <= 0 || d.H <= 0 || d.W <= 0 || d.P <= 0 || d.D <= 0)
    return set_error(DCV_ERR_INVALID, "embed: empty dims");
  if (d.H % d.P || d.W % d.P) return set_error(DCV_ERR_UNSUPPORTED, "embed: image size must be a multiple of the patch size");
  if ((d.P * d.P) % 8 || d.D % 64) return set_error(DCV_ERR_UNSUPPORTED, "embed: P*P %% 8 and D %% 64 required");
  return 0;
}

int embed_fwd(const dcv_embed_dims& d, const dcv_embed_cfg& cfg, const dcv_embed_params& p, const void* x,
              const int* idx, const int* gid, const dcv_embed_acts& a, cudaStream_t st) {
  DCV_TRY(check_embed(d));
  if (!x || !gid || !a.patches || !a.wsplit || !a.addend || !a.tokens || !a.extra)
    return set_error(DCV_ERR_INVALID, "embed_fwd: null pointer");
  const int N = (d.H / d.P) * (d.W / d.P), T = d.Cs * N, K = d.P * d.P, D = d.D;
  const bool fused = embed_fused_ok(d, cfg.x_is_u8);
  // DCS gather + unfold (dichavit.py:210, :377) -- inside the fused kernel when the shape allows it
  if (!fused)
    DCV_TRY(im2col_gather(x, cfg.x_is_u8, p.pix_mean, p.pix_inv_std, idx, a.patches, d.B, d.C, d.Cs, d.H, d.W, d.P, st));
  DCV_TRY(split_weight(p.proj_w, a.wsplit, D, K, st));
  // positional embedding of the patches: raw or bicubic-resampled (dichavit.py:529-552)
  const float* pos_patch = p.pos + D;
  if (p.pos_map) {
    DCV_TRY(sgemm_small(p.pos_map, N, 0, p.pos + D, D, 0, a.pos_patch, D, nullptr, 0, N, D, N, st));
    pos_patch = a.pos_patch;
  }
  DCV_TRY(embed_addend(p.proj_b, p.chan_embed, gid, pos_patch, p.cls, p.pos, a.addend, a.tokens, d.B, d.Cs, N, D, st));
  // tokens[b, 1 + t, :] = patches * W^T + addend[t]   (conv + bias + channel token + pos, :377,:409-411,:565)
  // (3K-wide split-precision operands: [hi|lo|hi] x [Whi|Whi|Wlo]^T, fp32 accumulation in TMEM)
  const bool tdl_on = cfg.lambda_tdl > 0.f, cdl_on = cfg.lambda_cdl > 0.f;
  if (fused) {
    // one kernel: TMA from the fp32 image -> bf16 hi/lo split in shared memory -> tcgen05 -> tokens + the per-(image,
    // channel) sums of normalised tokens TDL needs (embed_fused.cu)
    DCV_TRY(embed_fused_fwd(d, x, idx, a.wsplit, p.proj_b, a.addend, a.tokens, a.patches, a.S, a.Q, a.rnorm, tdl_on ? 1 : 0,
                            st));
    if (tdl_on)
      DCV_TRY(tdl_finish(a.S, a.Q, a.S_all, a.loss_b, a.coef_pos, a.coef_neg, a.tdl, d.B, d.Cs, N, D, cfg.gamma_s,
                         cfg.gamma_d, cfg.reverse_pos_pairs, cfg.use_square, st));
  } else {
    DCV_TRY(gemm_nt(a.patches, 3 * K, a.wsplit, 3 * K, d.B * T, D, 3 * K, EPI_EMBED, nullptr, a.tokens, nullptr, nullptr,
                    nullptr, D, false, st, T, T + 1, a.addend));
    if (tdl_on)
      DCV_TRY(tdl_fwd(a.tokens, a.addend, p.proj_b, a.S, a.Q, a.rnorm, a.S_all, a.loss_b, a.coef_pos, a.coef_neg, a.tdl,
                      d.B, d.Cs, N, D, cfg.gamma_s, cfg.gamma_d, cfg.reverse_pos_pairs, cfg.use_square, st));
  }
  if (cdl_on) {
    if (!p.proxies) return set_error(DCV_ERR_INVALID, "embed_fwd: CDL on but proxies == NULL");
    DCV_TRY(cdl_fwd(p.chan_embed, p.proxies, gid, cfg.cdl_scale, a.cdl, a.cdl_dE, a.cdl_dP, d.Cs, D, st));
  }
  DCV_TRY(extra_loss(tdl_on ? a.tdl : nullptr, cdl_on ? a.cdl : nullptr, cfg.lambda_tdl, cfg.lambda_cdl, a.extra, st));
  return 0;
}

int embed_bwd(const dcv_embed_dims& d, const dcv_embed_cfg& cfg, const dcv_embed_params& p, const int* gid,
              const dcv_embed_acts& a, const dcv_embed_grads& g, const dcv_embed_ws& ws, const float* G,
              const float* d_extra, cudaStream_t st) {
  DCV_TRY(check_embed(d));
  if (!G || !gid || !ws.dY || !ws.R) return set_error(DCV_ERR_INVALID, "embed_bwd: null pointer");
  const int N = (d.H / d.P) * (d.W / d.P), T = d.Cs * N, K = d.P * d.P, D = d.D, M = d.B * T;
  const bool tdl_on = cfg.lambda_tdl > 0.f && d_extra != nullptr, cdl_on = cfg.lambda_cdl > 0.f && d_extra != nullptr;
  // Two independent chains off the token gradient G: (1) dY -> conv weight / bias gradients, (2) batch sum -> cls / pos /
  // channel-token gradients (-> CDL's share of the channel-token gradient, same buffers, same stream).  Chain (2) is
  // five short latency-bound launches: it goes to the side stream (see side_branch() in host.h) when there is one.
  SideBranch* sb = side_branch(st, M);
  const cudaStream_t w = sb ? sb->s : st;
  if (sb) DCV_CUDA(side_fork(sb, st));
  // dY = G[:, 1:] + lambda_tdl * d_extra * dTDL/dY   (Appendix C of SURVEY.md)
  DCV_TRY(embed_bwd_dy(G, a.tokens, a.addend, p.proj_b, a.rnorm, tdl_on ? a.S : nullptr, a.S_all, a.coef_pos, a.coef_neg,
                       d_extra, tdl_on ? cfg.lambda_tdl : 0.f, ws.dY, d.B, d.Cs, N, D, st));
  // conv weight / bias gradients: dW[D, P*P] += dY^T patches ; db += colsum(dY)
  DCV_TRY(gemm_tn(ws.dY, D, a.patches, 3 * K, M, D, K, g.proj_w, K, 1, 0, st));  // hi part of the patches
  DCV_TRY(colsum_bf16(ws.dY, g.proj_b, M, D, D, st));
  // cls / pos / channel-token gradients from the batch-summed token gradient
  float* pos_patch_grad = p.pos_map ? ws.dpos_patch : g.pos + D;
  DCV_TRY(embed_param_grads(G, ws.R, gid, g.cls, g.pos, g.chan_embed, pos_patch_grad, p.pos_map ? 0 : 1, d.B, d.Cs, N, D, w));
  if (p.pos_map)  // d pos[1:] += pos_map^T dpos_patch
    DCV_TRY(sgemm_small(p.pos_map, N, 1, ws.dpos_patch, D, 0, g.pos + D, D, nullptr, 1, N, D, N, w));
  if (cdl_on) DCV_TRY(cdl_bwd(a.cdl_dE, a.cdl_dP, gid, d_extra, cfg.lambda_cdl, g.chan_embed, g.proxies, d.Cs, D, w));
  if (sb) {
    DCV_CUDA(side_mark(sb, 0));
    DCV_CUDA(side_join(sb, 0, st));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// head
// ---------------------------------------------------------------------------------------------
int head_fwd(const float* x_last, int B, int L, int D, const float* norm_w, const float* norm_b, float* feat,
             float* mean, float* rstd, const float* head_w, const float* head_b, float* logits, int num_classes,
             cudaStream_t st) {
  if (!x_last || !feat || !mean || !rstd) return set_error(DCV_ERR_INVALID, "head_fwd: null pointer");
  DCV_TRY(cls_ln_fwd(x_last, static_cast<long long>(L) * D, norm_w, norm_b, feat, mean, rstd, B, D, 1e-6f, st));
  if (head_w) {
    if (!logits || num_classes <= 0) return set_error(DCV_ERR_INVALID, "head_fwd: logits / num_classes missing");
    DCV_TRY(sgemm_small(feat, D, 0, head_w, D, 1, logits, num_classes, head_b, 0, B, num_classes, D, st));
  }
  return 0;
}

int head_bwd(const float* d_out, const float* x_last, int B, int L, int D, const float* norm_w, const float* feat,
             const float* mean, const float* rstd, const float* head_w, int num_classes, float* dfeat_ws, float* dres,
             void* dres_bf16, float* g_norm_w, float* g_norm_b, float* g_head_w, float* g_head_b, float* dbias_last,
             cudaStream_t st) {
  if (!d_out || !x_last || !dres || !dres_bf16) return set_error(DCV_ERR_INVALID, "head_bwd: null pointer");
  if (head_w && !dfeat_ws) return set_error(DCV_ERR_INVALID, "head_bwd: dfeat workspace missing");
  const size_t n = static_cast<size_t>(B) * L * D;
  // the two full-size clears run beside the head's small GEMMs (side stream), joined before the CLS rows are written
  SideBranch* sb = side_branch(st, B * L);
  const cudaStream_t w = sb ? sb->s : st;
  if (sb) DCV_CUDA(side_fork(sb, st));
  DCV_CUDA(cudaMemsetAsync(dres, 0, n * sizeof(float), w));
  DCV_CUDA(cudaMemsetAsync(dres_bf16, 0, n * 2, w));
  if (sb) DCV_CUDA(side_mark(sb, 0));
  const float* dfeat = d_out;
  if (head_w) {
    DCV_TRY(sgemm_small(d_out, num_classes, 0, head_w, D, 0, dfeat_ws, D, nullptr, 0, B, D, num_classes, st));
    if (g_head_w) DCV_TRY(sgemm_small(d_out, num_classes, 1, feat, D, 0, g_head_w, D, nullptr, 1, num_classes, D, B, st));
    if (g_head_b) DCV_TRY(colsum_f32(d_out, g_head_b, B, num_classes, num_classes, st));
    dfeat = dfeat_ws;
  }
  if (sb) DCV_CUDA(side_join(sb, 0, st));
  DCV_TRY(cls_ln_bwd(dfeat, x_last, static_cast<long long>(L) * D, mean, rstd, norm_w, dres, dres_bf16, g_norm_w, g_norm_b,
                     dbias_last, B, D, st));
  return 0;
}

}  // namespace dcv
