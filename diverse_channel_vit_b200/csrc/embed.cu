// HBM-bound kernels around the patch-embedding GEMM of the DiChaViT hot path
// (reference models/dichavit.py:110-417 PatchEmbedPerChannel.forward and :554-565
// prepare_tokens), the Token Diversification Loss in closed form (models/loss_fn.py:24-59)
// and the Channel Diversification Loss (models/loss_fn.py:7-21), each with its backward.
//
// Token layout everywhere: tokens fp32 [B, L = 1 + C'*N, D]; row 0 = CLS, row 1 + c*N + p =
// patch p (= hp * (W/P) + wp) of the c-th SAMPLED channel (sampled order, not sorted).
#include <algorithm>

#include "common.cuh"
#include "host.h"

namespace dcv {

// ---------------------------------------------------------------------------------
// DCS gather + im2col: x fp32 [B, C, H, W], idx int32 [C'] -> patches bf16 [B*C'*N, 3*P*P]
// (row = (b, c', hp, wp), column k = ph * P + pw).   reference dichavit.py:210 (x[:, idx]) and
// the unfold implied by Conv3d(1, D, (1,P,P), stride (1,P,P)) at :77-82,377.
// Each fp32 pixel is split into two bf16 (x = hi + lo, 16 mantissa bits kept); a row holds
// [hi | lo | hi] so that ONE bf16 GEMM against the split weight [Whi | Whi | Wlo] (below) yields
// hi*Whi + lo*Whi + hi*Wlo = x*W to ~2^-16: the reference projects in fp32 and the TDL scalar
// (tolerance 1e-3) is a difference of large sums of these outputs.
// One CTA per (b, c', hp) strip of P image rows: reads are full 128-byte lines along W,
// writes are 8-byte pieces that fill whole 32-byte sectors.
// ---------------------------------------------------------------------------------
// U8: x holds raw uint8 pixels and the loader's per-channel standardisation (x - mean[c]) / std[c] (reference
// datasets/dataset_utils.py:44, jump_cp_transforms.py:119-121) is applied here, on the device: the H2D copy shrinks 4x.
template <bool U8>
__global__ void __launch_bounds__(256)
im2col_gather_kernel(const void* __restrict__ xv, const int* __restrict__ idx, __nv_bfloat16* __restrict__ patches,
                     const float* __restrict__ pix_mean, const float* __restrict__ pix_inv_std, int C, int Cs, int H,
                     int W, int P) {
  const int hp_count = H / P, wp_count = W / P;
  int strip = blockIdx.x;
  const int hp = strip % hp_count;
  strip /= hp_count;
  const int cs = strip % Cs;
  const int b = strip / Cs;
  const int c_src = idx ? __ldg(idx + cs) : cs;
  const size_t src_off = ((static_cast<size_t>(b) * C + c_src) * H + static_cast<size_t>(hp) * P) * W;
  const float* src = reinterpret_cast<const float*>(xv) + src_off;
  const uint8_t* src8 = reinterpret_cast<const uint8_t*>(xv) + src_off;
  float mu = 0.f, is = 1.f;
  if (U8 && pix_mean != nullptr) {
    mu = __ldg(pix_mean + c_src);
    is = __ldg(pix_inv_std + c_src);
  }
  const int KK = P * P;
  __nv_bfloat16* dst = patches + ((static_cast<size_t>(b) * Cs + cs) * hp_count + hp) * wp_count * (3 * KK);
  const int w4 = W >> 2;
  for (int i = threadIdx.x; i < P * w4; i += blockDim.x) {
    const int ph = i / w4, wq = i - ph * w4;
    float4 v;
    if (U8) {
      const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(src8 + static_cast<size_t>(ph) * W) + wq);
      v = make_float4((u.x - mu) * is, (u.y - mu) * is, (u.z - mu) * is, (u.w - mu) * is);
    } else {
      v = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(ph) * W) + wq);
    }
    const int w = wq * 4;
    const int wp = w / P, pw = w - wp * P;
    uint2 hi, lo;
    hi.x = pack_bf16(v.x, v.y);
    hi.y = pack_bf16(v.z, v.w);
    const float2 h01 = unpack_bf16(hi.x), h23 = unpack_bf16(hi.y);
    lo.x = pack_bf16(v.x - h01.x, v.y - h01.y);
    lo.y = pack_bf16(v.z - h23.x, v.w - h23.y);
    __nv_bfloat16* row = dst + static_cast<size_t>(wp) * (3 * KK) + ph * P + pw;
    *reinterpret_cast<uint2*>(row) = hi;
    *reinterpret_cast<uint2*>(row + KK) = lo;
    *reinterpret_cast<uint2*>(row + 2 * KK) = hi;
  }
}

// conv weight fp32 [D, K] -> bf16 [D, 3K] = [Whi | Whi | Wlo]  (see im2col_gather_kernel)
__global__ void split_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ ws, int D, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D * K) return;
  const int d = i / K, k = i - d * K;
  const float v = w[i];
  const __nv_bfloat16 hi = __float2bfloat16(v);
  const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
  __nv_bfloat16* row = ws + static_cast<size_t>(d) * (3 * K);
  row[k] = hi;
  row[K + k] = hi;
  row[2 * K + k] = lo;
}

// ---------------------------------------------------------------------------------
// addend[t = c*N + p, :] = conv bias + channel_embed[gid[c]] + pos_patch[p]   (fp32 [T, D])
// and the CLS rows tokens[b, 0, :] = cls_token + pos_embed[0].
// reference dichavit.py:409-411 (+channel_embed), :561-565 (cat CLS, + pos).
// ---------------------------------------------------------------------------------
__global__ void embed_addend_kernel(const float* __restrict__ bias, const float* __restrict__ chan_embed,
                                    const int* __restrict__ gid, const float* __restrict__ pos_patch,
                                    const float* __restrict__ cls, const float* __restrict__ pos0,
                                    float* __restrict__ addend, float* __restrict__ tokens, int B, int Cs, int N,
                                    int D) {
  const int d4 = D >> 2;
  const long long T = static_cast<long long>(Cs) * N;
  const long long total = (T + B) * d4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % d4);
    const long long r = i / d4;
    if (r < T) {
      const int c = static_cast<int>(r / N), p = static_cast<int>(r - static_cast<long long>(c) * N);
      const float4 a = __ldg(reinterpret_cast<const float4*>(bias) + col);
      const float4 e = __ldg(reinterpret_cast<const float4*>(chan_embed + static_cast<size_t>(__ldg(gid + c)) * D) + col);
      const float4 q = __ldg(reinterpret_cast<const float4*>(pos_patch + static_cast<size_t>(p) * D) + col);
      reinterpret_cast<float4*>(addend + r * D)[col] =
          make_float4(a.x + e.x + q.x, a.y + e.y + q.y, a.z + e.z + q.z, a.w + e.w + q.w);
    } else {
      const long long b = r - T;
      const float4 a = __ldg(reinterpret_cast<const float4*>(cls) + col);
      const float4 q = __ldg(reinterpret_cast<const float4*>(pos0) + col);
      reinterpret_cast<float4*>(tokens + b * (T + 1) * D)[col] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
    }
  }
}

// ---------------------------------------------------------------------------------
// TDL, closed form.  With f_i = normalize(y_i) (y = conv output + bias, BEFORE the
// channel/pos add; F.normalize eps 1e-12) and S_c = sum_{i in channel c} f_i:
//   pos_sum = sum_c |S_c|^2 - sum_i |f_i|^2        neg_sum = |sum_c S_c|^2 - sum_c |S_c|^2
// which equals the masked Gram sums of models/loss_fn.py:36-48 without the B x T x T matrix.
//
// tdl_sum_kernel: one CTA -- or one thread-block cluster of `split` CTAs, each taking a slice of the tokens, when
// B*C' alone would not fill the 148 SMs -- per (b, c); warps stride over the channel's N tokens; y is recovered
// from the stored token as tokens - addend + bias (the GEMM epilogue added addend).  The cluster's partial sums are
// combined by its rank-0 CTA through distributed shared memory in rank order: no atomics, deterministic.
//   S [B, C', D] fp32, Q [B, C'] = sum_i |f_i|^2, rnorm [B, T] = 1 / max(|y_i|, eps)
// ---------------------------------------------------------------------------------
constexpr int kTdlWarps = 8;

template <int NV>
__global__ void __launch_bounds__(kTdlWarps * 32)
tdl_sum_kernel(const float* __restrict__ tokens, const float* __restrict__ addend, const float* __restrict__ bias,
               float* __restrict__ S, float* __restrict__ Q, float* __restrict__ rnorm, int Cs, int N, int D,
               int split) {
  extern __shared__ __align__(16) float red[];  // [kTdlWarps][D] + [kTdlWarps]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 2;
  const int bc = blockIdx.x / split;
  const int rank = blockIdx.x - bc * split;  // == %cluster_ctarank
  const int per = (N + split - 1) / split;
  const int p_lo = rank * per, p_hi = min(N, p_lo + per);
  const int b = bc / Cs, c = bc - b * Cs;
  const int T = Cs * N;
  float4 bi[NV], acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    bi[k] = v < nvec ? __ldg(reinterpret_cast<const float4*>(bias) + v) : make_float4(0, 0, 0, 0);
    acc[k] = make_float4(0, 0, 0, 0);
  }
  float q = 0.f;
  // software pipeline: the next token's loads are in flight while this one is normalised and accumulated
  auto load_row = [&](int p, float4 (&tv)[NV], float4 (&av)[NV]) {
    const int t = c * N + p;
    const float4* tr = reinterpret_cast<const float4*>(tokens + (static_cast<size_t>(b) * (T + 1) + 1 + t) * D);
    const float4* ar = reinterpret_cast<const float4*>(addend + static_cast<size_t>(t) * D);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      tv[k] = v < nvec ? __ldcs(tr + v) : make_float4(0, 0, 0, 0);
      av[k] = v < nvec ? __ldg(ar + v) : make_float4(0, 0, 0, 0);
    }
  };
  float4 tv[NV], av[NV], tn[NV], an[NV];
  if (p_lo + warp < p_hi) load_row(p_lo + warp, tv, av);
  for (int p = p_lo + warp; p < p_hi; p += kTdlWarps) {
    const int t = c * N + p;
    if (p + kTdlWarps < p_hi) load_row(p + kTdlWarps, tn, an);
    float4 y[NV];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nvec) {
        y[k] = make_float4(tv[k].x - av[k].x + bi[k].x, tv[k].y - av[k].y + bi[k].y, tv[k].z - av[k].z + bi[k].z,
                           tv[k].w - av[k].w + bi[k].w);
        ss += (y[k].x * y[k].x + y[k].y * y[k].y) + (y[k].z * y[k].z + y[k].w * y[k].w);
      } else {
        y[k] = make_float4(0, 0, 0, 0);
      }
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      acc[k].x += y[k].x * inv; acc[k].y += y[k].y * inv; acc[k].z += y[k].z * inv; acc[k].w += y[k].w * inv;
      tv[k] = tn[k];
      av[k] = an[k];
    }
    q += ss * inv * inv;
    if (lane == 0) rnorm[static_cast<size_t>(b) * T + t] = inv;
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nvec) reinterpret_cast<float4*>(red + warp * D)[v] = acc[k];
  }
  if (lane == 0) red[kTdlWarps * D + warp] = q;
  __syncthreads();
  if (split == 1) {
    for (int col = threadIdx.x; col < D; col += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kTdlWarps; ++w) t += red[w * D + col];
      S[static_cast<size_t>(bc) * D + col] = t;
    }
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kTdlWarps; ++w) t += red[kTdlWarps * D + w];
      Q[bc] = t;
    }
    return;
  }
  // cluster path: every CTA folds its warps into row 0 of `red`, rank 0 then adds the peers' rows in rank order
  float tcol[4];  // D <= 4 * blockDim.x
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int col = threadIdx.x + k * blockDim.x;
    float t = 0.f;
    if (col < D) {
#pragma unroll
      for (int w = 0; w < kTdlWarps; ++w) t += red[w * D + col];
    }
    tcol[k] = t;
  }
  __syncthreads();  // every warp row has been read before row 0 is overwritten
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int col = threadIdx.x + k * blockDim.x;
    if (col < D) red[col] = tcol[k];
  }
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kTdlWarps; ++w) t += red[kTdlWarps * D + w];
    red[kTdlWarps * D] = t;
  }
  cluster_sync_all();
  if (rank == 0) {
    for (int col = threadIdx.x; col < D; col += blockDim.x) {
      float t = red[col];
      for (int rk = 1; rk < split; ++rk) t += ld_dsmem_f32(&red[col], rk);
      S[static_cast<size_t>(bc) * D + col] = t;
    }
    if (threadIdx.x == 0) {
      float t = red[kTdlWarps * D];
      for (int rk = 1; rk < split; ++rk) t += ld_dsmem_f32(&red[kTdlWarps * D], rk);
      Q[bc] = t;
    }
  }
  cluster_sync_all();  // peers keep their shared memory alive until rank 0 has read it
}

// block-wide sum of one float per thread (blockDim.x <= 1024, multiple of 32)
__device__ __forceinline__ float block_sum(float v, float* scratch /*[32]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) t += scratch[w];
  return t;
}

// tdl_pair_kernel: one CTA per image: S_all[b,:] = sum_c S[b,c,:]; pos/neg means; per-image loss
// (models/loss_fn.py:44-58) and the two backward coefficients dLoss/dpos_sum, dLoss/dneg_sum
// (already divided by B for the batch mean at :59).
struct TdlFlags {
  float gamma_s, gamma_d;
  int reverse_pos_pairs, use_square;
};

__global__ void __launch_bounds__(128)
tdl_pair_kernel(const float* __restrict__ S, const float* __restrict__ Q, float* __restrict__ S_all,
                float* __restrict__ loss_b, float* __restrict__ coef_pos, float* __restrict__ coef_neg, int B,
                int Cs, int N, int D, TdlFlags f) {
  __shared__ float scratch[32];
  const int b = blockIdx.x;
  float a = 0.f, e = 0.f;
  for (int col = threadIdx.x; col < D; col += blockDim.x) {
    float sall = 0.f;
    for (int c = 0; c < Cs; ++c) {
      const float s = S[(static_cast<size_t>(b) * Cs + c) * D + col];
      a += s * s;
      sall += s;
    }
    S_all[static_cast<size_t>(b) * D + col] = sall;
    e += sall * sall;
  }
  a = block_sum(a, scratch);
  e = block_sum(e, scratch);
  if (threadIdx.x == 0) {
    float qb = 0.f;
    for (int c = 0; c < Cs; ++c) qb += Q[b * Cs + c];
    const float T = static_cast<float>(Cs) * static_cast<float>(N);
    const float pos_n = static_cast<float>(Cs) * static_cast<float>(N) * static_cast<float>(N - 1) + 1e-6f;
    const float neg_n = T * T - static_cast<float>(Cs) * static_cast<float>(N) * static_cast<float>(N) + 1e-6f;
    float pos = (a - qb) / pos_n;
    float neg = (e - a) / neg_n;
    float dneg = f.gamma_d, dpos;
    if (f.use_square) {
      dneg = f.gamma_d * 2.f * neg;
      neg = neg * neg;
    }
    float loss;
    if (f.reverse_pos_pairs) {
      dpos = f.gamma_s;
      if (f.use_square) {
        dpos = f.gamma_s * 2.f * pos;
        pos = pos * pos;
      }
      loss = f.gamma_s * pos + f.gamma_d * neg;
    } else {
      dpos = -f.gamma_s;
      loss = f.gamma_s * (1.0f - pos) + f.gamma_d * neg;
    }
    loss_b[b] = loss;
    coef_pos[b] = dpos / pos_n / static_cast<float>(B);
    coef_neg[b] = dneg / neg_n / static_cast<float>(B);
  }
}

// out[0] = scale * sum(v[0..n))   (single CTA; deterministic order)
__global__ void __launch_bounds__(256) reduce_sum_kernel(const float* __restrict__ v, int n, float scale,
                                                         float* __restrict__ out) {
  __shared__ float scratch[32];
  float t = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += v[i];
  t = block_sum(t, scratch);
  if (threadIdx.x == 0) out[0] = t * scale;
}

// ---------------------------------------------------------------------------------
// Backward of the token assembly + TDL: per token (one warp each)
//   dY[b,t,:] = G[b,1+t,:] + k * dTDL/dy          -> bf16 [B*T, D], A operand of the conv wgrad GEMM
// with k = lambda_tdl * d(extra_loss) read from device memory (no host sync), and
//   dTDL/df_i = 2 coef_pos[b] (S_c - f_i) + 2 coef_neg[b] (S_all - S_c),  dy = (g - f (f.g)) / |y|.
// ---------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
embed_bwd_dy_kernel(const float* __restrict__ G, const float* __restrict__ tokens, const float* __restrict__ addend,
                    const float* __restrict__ bias, const float* __restrict__ rnorm, const float* __restrict__ S,
                    const float* __restrict__ S_all, const float* __restrict__ coef_pos,
                    const float* __restrict__ coef_neg, const float* __restrict__ d_extra, float lambda_tdl,
                    __nv_bfloat16* __restrict__ dY, int B, int Cs, int N, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 2;
  const int T = Cs * N;
  const long long rows = static_cast<long long>(B) * T;
  const bool with_tdl = lambda_tdl != 0.f && S != nullptr;
  const float kk = with_tdl ? lambda_tdl * __ldg(d_extra) : 0.f;
  for (long long r = static_cast<long long>(blockIdx.x) * 8 + warp; r < rows; r += static_cast<long long>(gridDim.x) * 8) {
    const int b = static_cast<int>(r / T), t = static_cast<int>(r - static_cast<long long>(b) * T);
    const size_t tok_off = (static_cast<size_t>(b) * (T + 1) + 1 + t) * D;
    const float4* gr = reinterpret_cast<const float4*>(G + tok_off);
    float4 g[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      g[k] = v < nvec ? gr[v] : make_float4(0, 0, 0, 0);
    }
    if (with_tdl) {
      const int c = t / N;
      const float4* tr = reinterpret_cast<const float4*>(tokens + tok_off);
      const float4* ar = reinterpret_cast<const float4*>(addend + static_cast<size_t>(t) * D);
      const float4* sc = reinterpret_cast<const float4*>(S + (static_cast<size_t>(b) * Cs + c) * D);
      const float4* sa = reinterpret_cast<const float4*>(S_all + static_cast<size_t>(b) * D);
      const float inv = rnorm[r];
      const float cp = 2.f * kk * coef_pos[b], cn = 2.f * kk * coef_neg[b];
      float4 f[NV], gf[NV];
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int v = lane + 32 * k;
        if (v < nvec) {
          const float4 tv = tr[v], av = __ldg(ar + v), bv = __ldg(reinterpret_cast<const float4*>(bias) + v);
          const float4 s1 = __ldg(sc + v), s2 = __ldg(sa + v);
          f[k] = make_float4((tv.x - av.x + bv.x) * inv, (tv.y - av.y + bv.y) * inv, (tv.z - av.z + bv.z) * inv,
                             (tv.w - av.w + bv.w) * inv);
          gf[k] = make_float4(cp * (s1.x - f[k].x) + cn * (s2.x - s1.x), cp * (s1.y - f[k].y) + cn * (s2.y - s1.y),
                              cp * (s1.z - f[k].z) + cn * (s2.z - s1.z), cp * (s1.w - f[k].w) + cn * (s2.w - s1.w));
          dot += (f[k].x * gf[k].x + f[k].y * gf[k].y) + (f[k].z * gf[k].z + f[k].w * gf[k].w);
        } else {
          f[k] = gf[k] = make_float4(0, 0, 0, 0);
        }
      }
      dot = warp_sum(dot);
      if (inv >= 1e12f) dot = 0.f;  // |y| <= eps: F.normalize divides by the constant eps
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        g[k].x += inv * (gf[k].x - f[k].x * dot);
        g[k].y += inv * (gf[k].y - f[k].y * dot);
        g[k].z += inv * (gf[k].z - f[k].z * dot);
        g[k].w += inv * (gf[k].w - f[k].w * dot);
      }
    }
    uint2* dr = reinterpret_cast<uint2*>(dY + static_cast<size_t>(r) * D);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nvec) {
        uint2 o;
        o.x = pack_bf16(g[k].x, g[k].y);
        o.y = pack_bf16(g[k].z, g[k].w);
        dr[v] = o;
      }
    }
  }
}

// R[l, :] = sum_b G[b, l, :]   (G fp32 [B, L, D]); one thread per float4 column of one token row
__global__ void __launch_bounds__(256)
batch_sum_kernel(const float* __restrict__ G, float* __restrict__ R, int B, long long LD4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= LD4) return;
  const float4* g = reinterpret_cast<const float4*>(G) + i;
  float4 a = make_float4(0, 0, 0, 0);
  int b = 0;
  for (; b + 4 <= B; b += 4) {
    const float4 v0 = g[(b + 0) * LD4], v1 = g[(b + 1) * LD4], v2 = g[(b + 2) * LD4], v3 = g[(b + 3) * LD4];
    a.x += (v0.x + v1.x) + (v2.x + v3.x);
    a.y += (v0.y + v1.y) + (v2.y + v3.y);
    a.z += (v0.z + v1.z) + (v2.z + v3.z);
    a.w += (v0.w + v1.w) + (v2.w + v3.w);
  }
  for (; b < B; ++b) {
    const float4 v = g[b * LD4];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  reinterpret_cast<float4*>(R)[i] = a;
}

// From R [L, D]: d cls_token += R[0]; d pos_embed[0] += R[0];
//   d channel_embed[gid[c]] += sum_p R[1 + c*N + p];  dpos_patch[p] (=/+=) sum_c R[1 + c*N + p]
// grid.x = 1 + C' + N tasks; blockDim = 4 * (D / 4): thread = (float4 column, one of 4 slices of the task's rows), the
// four partial sums meet in shared memory.  (The first version walked the N = 196 rows of a channel task with one
// load in flight per thread and three columns per thread: 22 us for 0.6 MB.)
__global__ void __launch_bounds__(1024)
embed_param_grads_kernel(const float* __restrict__ R, const int* __restrict__ gid, float* __restrict__ d_cls,
                         float* __restrict__ d_pos0, float* __restrict__ d_chan_embed, float* __restrict__ dpos_patch,
                         int accumulate_pos, int Cs, int N, int D) {
  extern __shared__ float4 pg_part[];  // [4][D / 4]
  const int nv = D >> 2;
  const int c4 = threadIdx.x % nv, sl = threadIdx.x / nv;  // sl in 0..3
  const int task = blockIdx.x;
  // rows of R this task sums: first + i * step, i in [0, n)
  int first, step, n;
  if (task == 0) {
    first = 0; step = 1; n = 1;
  } else if (task <= Cs) {
    first = 1 + (task - 1) * N; step = 1; n = N;
  } else {
    first = 1 + (task - 1 - Cs); step = N; n = Cs;
  }
  const float4* R4 = reinterpret_cast<const float4*>(R);
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int i = sl; i < n; i += 4) {
    const float4 v = R4[static_cast<size_t>(first + i * step) * nv + c4];
    t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
  }
  pg_part[sl * nv + c4] = t;
  __syncthreads();
  if (sl != 0) return;
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    const float4 v = pg_part[k * nv + c4];
    t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
  }
  if (task == 0) {
    float4* a = reinterpret_cast<float4*>(d_cls) + c4;
    float4* b = reinterpret_cast<float4*>(d_pos0) + c4;
    float4 va = *a, vb = *b;
    va.x += t.x; va.y += t.y; va.z += t.z; va.w += t.w;
    vb.x += t.x; vb.y += t.y; vb.z += t.z; vb.w += t.w;
    *a = va;
    *b = vb;
  } else if (task <= Cs) {
    if (d_chan_embed) {
      float4* a = reinterpret_cast<float4*>(d_chan_embed + static_cast<size_t>(__ldg(gid + task - 1)) * D) + c4;
      float4 va = *a;
      va.x += t.x; va.y += t.y; va.z += t.z; va.w += t.w;
      *a = va;
    }
  } else {
    float4* a = reinterpret_cast<float4*>(dpos_patch + static_cast<size_t>(task - 1 - Cs) * D) + c4;
    if (accumulate_pos) {
      const float4 va = *a;
      t.x += va.x; t.y += va.y; t.z += va.z; t.w += va.w;
    }
    *a = t;
  }
}

// ---------------------------------------------------------------------------------
// CDL: proxy_loss(channel_emb_proxies[gid], channel_embed[gid], eye(C'), scale)  (loss_fn.py:7-21)
//   e = normalize(E), p = normalize(P); z_ij = -scale^2 |e_i - p_j|^2; loss = mean_i -log softmax(z_i)_i
// Single CTA (C' <= 32 rows); forward also produces the unscaled gradients dE, dP [C', D];
// cdl_bwd scatters coef * dE/dP into the parameter gradients (rows gid[c]).
// ---------------------------------------------------------------------------------
constexpr int kCdlMaxC = 32;

__global__ void __launch_bounds__(256)
cdl_fwd_kernel(const float* __restrict__ chan_embed, const float* __restrict__ proxies, const int* __restrict__ gid,
               float scale, float* __restrict__ loss, float* __restrict__ dE, float* __restrict__ dP, int Cs, int D) {
  extern __shared__ float sm[];  // e[Cs][D], p[Cs][D]
  __shared__ float en[kCdlMaxC], pn[kCdlMaxC], z[kCdlMaxC][kCdlMaxC + 1], q[kCdlMaxC][kCdlMaxC + 1];
  __shared__ float row_loss[kCdlMaxC];
  float* e = sm;
  float* p = sm + static_cast<size_t>(Cs) * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // load + normalise rows (one warp per row)
  for (int r = warp; r < 2 * Cs; r += nw) {
    const int c = r % Cs;
    const float* src = (r < Cs ? chan_embed : proxies) + static_cast<size_t>(__ldg(gid + c)) * D;
    float* dst = (r < Cs ? e : p) + static_cast<size_t>(c) * D;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float v = src[d];
      dst[d] = v;
      ss += v * v;
    }
    ss = warp_sum(ss);
    const float nrm = fmaxf(sqrtf(ss), 1e-12f);
    for (int d = lane; d < D; d += 32) dst[d] /= nrm;
    if (lane == 0) (r < Cs ? en : pn)[c] = nrm;
  }
  __syncthreads();
  const float s2 = scale * scale;
  for (int ij = warp; ij < Cs * Cs; ij += nw) {
    const int i = ij / Cs, j = ij - i * Cs;
    float dd = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float df = e[i * D + d] - p[j * D + d];
      dd += df * df;
    }
    dd = warp_sum(dd);
    if (lane == 0) z[i][j] = -s2 * dd;
  }
  __syncthreads();
  if (threadIdx.x < Cs) {
    const int i = threadIdx.x;
    float mx = -INFINITY;
    for (int j = 0; j < Cs; ++j) mx = fmaxf(mx, z[i][j]);
    float sum = 0.f;
    for (int j = 0; j < Cs; ++j) sum += expf(z[i][j] - mx);
    const float lse = mx + logf(sum);
    row_loss[i] = lse - z[i][i];
    for (int j = 0; j < Cs; ++j) q[i][j] = (expf(z[i][j] - lse) - (i == j ? 1.f : 0.f)) / static_cast<float>(Cs);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < Cs; ++i) t += row_loss[i];
    loss[0] = t / static_cast<float>(Cs);
  }
  // gradients wrt the normalised rows, then through the normalisation (one warp per row)
  for (int r = warp; r < 2 * Cs; r += nw) {
    const bool is_e = r < Cs;
    const int c = r % Cs;
    const float* self = (is_e ? e : p) + static_cast<size_t>(c) * D;
    const float* other = is_e ? p : e;
    float* out = (is_e ? dE : dP) + static_cast<size_t>(c) * D;
    // dz/d self = -2 s2 (self - other_j)
    float dot = 0.f;
    for (int d = lane; d < D; d += 32) {
      float gsum = 0.f;
      for (int j = 0; j < Cs; ++j) {
        const float w = is_e ? q[c][j] : q[j][c];
        gsum += w * (-2.f * s2) * (self[d] - other[j * D + d]);
      }
      out[d] = gsum;
      dot += gsum * self[d];
    }
    dot = warp_sum(dot);
    const float nrm = (is_e ? en : pn)[c];
    for (int d = lane; d < D; d += 32) out[d] = (out[d] - self[d] * dot) / nrm;
  }
}

__global__ void cdl_bwd_kernel(const float* __restrict__ dE, const float* __restrict__ dP, const int* __restrict__ gid,
                               const float* __restrict__ d_extra, float lambda_cdl, float* __restrict__ g_chan_embed,
                               float* __restrict__ g_proxies, int Cs, int D) {
  const float k = lambda_cdl * __ldg(d_extra);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Cs * D; i += gridDim.x * blockDim.x) {
    const int c = i / D, d = i - c * D;
    const size_t o = static_cast<size_t>(__ldg(gid + c)) * D + d;
    if (g_chan_embed) g_chan_embed[o] += k * dE[i];
    if (g_proxies) g_proxies[o] += k * dP[i];
  }
}

// extra[0] = lambda_tdl * tdl + lambda_cdl * cdl   (dichavit.py:406-408)
__global__ void extra_loss_kernel(const float* tdl, const float* cdl, float lt, float lc, float* extra) {
  extra[0] = (tdl ? lt * tdl[0] : 0.f) + (cdl ? lc * cdl[0] : 0.f);
}

// ---------------------------------------------------------------------------------
// Small fp32 SIMT GEMM for the KB-sized contractions (bicubic pos-embed resample and its
// transpose, classifier head forward / dgrad / wgrad):
//   C[M,N] = (accumulate ? C : 0) + op(A)[M,K] * op(B)[K,N] (+ bias[N])
// op(A) = A (lda = K-stride) or A^T (A stored [K,M]); op(B) = B ([K,N]) or B^T (B stored [N,K]).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgemm_small_kernel(const float* __restrict__ A, int lda, int transA, const float* __restrict__ Bm, int ldb, int transB,
                   float* __restrict__ C, int ldc, const float* __restrict__ bias, int accumulate, int M, int N,
                   int K) {
  __shared__ float As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      // As[m][k]
      {
        const int m = transA ? m0 + tx : m0 + r, k = transA ? k0 + r : k0 + tx;
        float v = 0.f;
        if (m < M && k < K) v = transA ? A[static_cast<size_t>(k) * lda + m] : A[static_cast<size_t>(m) * lda + k];
        if (transA) As[tx][r] = v; else As[r][tx] = v;
      }
      // Bs[k][n]
      {
        const int k = transB ? k0 + tx : k0 + r, n = transB ? n0 + r : n0 + tx;
        float v = 0.f;
        if (k < K && n < N) v = transB ? Bm[static_cast<size_t>(n) * ldb + k] : Bm[static_cast<size_t>(k) * ldb + n];
        if (transB) Bs[tx][r] = v; else Bs[r][tx] = v;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float bv = Bs[k][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(As[ty + 8 * i][k], bv, acc[i]);
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n < N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty + 8 * i;
      if (m < M) {
        float v = acc[i] + (bias ? bias[n] : 0.f);
        float* dst = C + static_cast<size_t>(m) * ldc + n;
        *dst = accumulate ? *dst + v : v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// Final LayerNorm on the CLS rows only (reference dichavit.py:651-652 norms all rows, keeps row 0)
//   feat fp32 [B, D] = LN(x[b, 0, :]);  backward writes the CLS rows of the (zeroed) residual
//   gradient, its bf16 copy, dgamma/dbeta, and the column sum (bias gradient of the last fc2).
// One warp per image.
// ---------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(128)
cls_ln_fwd_kernel(const float* __restrict__ x, long long row_stride, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ feat, float* __restrict__ mean,
                  float* __restrict__ rstd, int B, int D, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= B) return;
  const int nvec = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(b) * row_stride);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    v[k] = i < nvec ? xr[i] : make_float4(0, 0, 0, 0);
    s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  const float mu = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < nvec) {
      const float a = v[k].x - mu, bq = v[k].y - mu, c = v[k].z - mu, d = v[k].w - mu;
      q += (a * a + bq * bq) + (c * c + d * d);
    }
  }
  const float rs = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
  float4* fr = reinterpret_cast<float4*>(feat + static_cast<size_t>(b) * D);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i), bt = __ldg(reinterpret_cast<const float4*>(beta) + i);
      fr[i] = make_float4((v[k].x - mu) * rs * g.x + bt.x, (v[k].y - mu) * rs * g.y + bt.y,
                          (v[k].z - mu) * rs * g.z + bt.z, (v[k].w - mu) * rs * g.w + bt.w);
    }
  }
  if (lane == 0) {
    mean[b] = mu;
    rstd[b] = rs;
  }
}

template <int NV>
__global__ void __launch_bounds__(128)
cls_ln_bwd_kernel(const float* __restrict__ dfeat, const float* __restrict__ x, long long row_stride,
                  const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                  float* __restrict__ dres, __nv_bfloat16* __restrict__ dres_bf16, float* __restrict__ dgamma,
                  float* __restrict__ dbeta, float* __restrict__ dxsum, int B, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= B) return;
  const int nvec = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(b) * row_stride);
  const float4* dyr = reinterpret_cast<const float4*>(dfeat + static_cast<size_t>(b) * D);
  const float mu = mean[b], rs = rstd[b];
  float4 xh[NV], gy[NV], dy[NV];
  float c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < nvec) {
      const float4 xv = xr[i], g = __ldg(reinterpret_cast<const float4*>(gamma) + i);
      dy[k] = dyr[i];
      xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      gy[k] = make_float4(dy[k].x * g.x, dy[k].y * g.y, dy[k].z * g.z, dy[k].w * g.w);
      c1 += (gy[k].x + gy[k].y) + (gy[k].z + gy[k].w);
      c2 += (gy[k].x * xh[k].x + gy[k].y * xh[k].y) + (gy[k].z * xh[k].z + gy[k].w * xh[k].w);
    } else {
      xh[k] = gy[k] = dy[k] = make_float4(0, 0, 0, 0);
    }
  }
  c1 = warp_sum(c1) / static_cast<float>(D);
  c2 = warp_sum(c2) / static_cast<float>(D);
  float4* dr = reinterpret_cast<float4*>(dres + static_cast<size_t>(b) * row_stride);
  uint2* db = reinterpret_cast<uint2*>(dres_bf16 + static_cast<size_t>(b) * row_stride);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < nvec) {
      const float4 o = make_float4(rs * (gy[k].x - c1 - xh[k].x * c2), rs * (gy[k].y - c1 - xh[k].y * c2),
                                   rs * (gy[k].z - c1 - xh[k].z * c2), rs * (gy[k].w - c1 - xh[k].w * c2));
      dr[i] = o;
      uint2 ob;
      ob.x = pack_bf16(o.x, o.y);
      ob.y = pack_bf16(o.z, o.w);
      db[i] = ob;
      const int col = i * 4;
      atomicAdd(dgamma + col + 0, dy[k].x * xh[k].x); atomicAdd(dgamma + col + 1, dy[k].y * xh[k].y);
      atomicAdd(dgamma + col + 2, dy[k].z * xh[k].z); atomicAdd(dgamma + col + 3, dy[k].w * xh[k].w);
      atomicAdd(dbeta + col + 0, dy[k].x); atomicAdd(dbeta + col + 1, dy[k].y);
      atomicAdd(dbeta + col + 2, dy[k].z); atomicAdd(dbeta + col + 3, dy[k].w);
      if (dxsum) {
        atomicAdd(dxsum + col + 0, o.x); atomicAdd(dxsum + col + 1, o.y);
        atomicAdd(dxsum + col + 2, o.z); atomicAdd(dxsum + col + 3, o.w);
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
#define DCV_NV_SWITCH(D, ...)                                                                 \
  switch (((D) + 127) / 128) {                                                                  \
    case 1: { constexpr int NV = 1; __VA_ARGS__; } break;                                              \
    case 2: { constexpr int NV = 2; __VA_ARGS__; } break;                                              \
    case 3: { constexpr int NV = 3; __VA_ARGS__; } break;                                              \
    case 6: { constexpr int NV = 6; __VA_ARGS__; } break;                                              \
    default: return set_error(DCV_ERR_UNSUPPORTED, "embed dim %d not instantiated (<=384 or 768)", (D)); \
  }

int im2col_gather(const void* x, int x_is_u8, const float* pix_mean, const float* pix_inv_std, const int* idx,
                  void* patches, int B, int C, int Cs, int H, int W, int P, cudaStream_t st) {
  if (B <= 0 || Cs <= 0 || C <= 0) return set_error(DCV_ERR_INVALID, "im2col: empty problem");
  if (H % P || W % P || P % 4 || W % 4) return set_error(DCV_ERR_UNSUPPORTED, "im2col: H, W multiples of P; P, W of 4");
  ProfScope prof(PT_IM2COL, st);
  const long long strips = static_cast<long long>(B) * Cs * (H / P);
  if ((pix_mean == nullptr) != (pix_inv_std == nullptr))
    return set_error(DCV_ERR_INVALID, "im2col: pix_mean and pix_inv_std must be given together");
  if (x_is_u8)
    im2col_gather_kernel<true><<<static_cast<unsigned>(strips), 256, 0, st>>>(
        x, idx, reinterpret_cast<__nv_bfloat16*>(patches), pix_mean, pix_inv_std, C, Cs, H, W, P);
  else
    im2col_gather_kernel<false><<<static_cast<unsigned>(strips), 256, 0, st>>>(
        x, idx, reinterpret_cast<__nv_bfloat16*>(patches), nullptr, nullptr, C, Cs, H, W, P);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int split_weight(const float* w, void* ws, int D, int K, cudaStream_t st) {
  ProfScope prof(PT_EMBED_MISC, st);
  split_weight_kernel<<<(D * K + 255) / 256, 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16*>(ws), D, K);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int embed_addend(const float* bias, const float* chan_embed, const int* gid, const float* pos_patch, const float* cls,
                 const float* pos0, float* addend, float* tokens, int B, int Cs, int N, int D, cudaStream_t st) {
  if (D % 4) return set_error(DCV_ERR_UNSUPPORTED, "embed_addend: D %% 4");
  ProfScope prof(PT_EMBED_MISC, st);
  const long long total = (static_cast<long long>(Cs) * N + B) * (D / 4);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
  embed_addend_kernel<<<blocks, 256, 0, st>>>(bias, chan_embed, gid, pos_patch, cls, pos0, addend, tokens, B, Cs, N, D);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int tdl_fwd(const float* tokens, const float* addend, const float* bias, float* S, float* Q, float* rnorm,
            float* S_all, float* loss_b, float* coef_pos, float* coef_neg, float* tdl_out, int B, int Cs, int N, int D,
            float gamma_s, float gamma_d, int reverse_pos_pairs, int use_square, cudaStream_t st) {
  if (B <= 0 || Cs <= 0 || N <= 0) return set_error(DCV_ERR_INVALID, "tdl_fwd: empty problem");
  if (D % 4) return set_error(DCV_ERR_UNSUPPORTED, "tdl_fwd: D %% 4");
  ProfScope prof(PT_TDL, st);
  const size_t smem = (static_cast<size_t>(kTdlWarps) * D + kTdlWarps) * sizeof(float);
  {
    // token-sliced clusters when B*C' CTAs alone leave SMs idle (JUMP-CP: 32 x 8 = 256 CTAs for 148 SMs)
    const int split = (B * Cs < 4 * num_sms() && N >= 64) ? 4 : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * Cs * split);
    cfg.blockDim = dim3(kTdlWarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = split;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = split > 1 ? 1 : 0;
    DCV_NV_SWITCH(D, DCV_CUDA(cudaLaunchKernelEx(&cfg, tdl_sum_kernel<NV>, tokens, addend, bias, S, Q, rnorm, Cs, N, D, split)));
  }
  count_launch();
  return tdl_finish(S, Q, S_all, loss_b, coef_pos, coef_neg, tdl_out, B, Cs, N, D, gamma_s, gamma_d, reverse_pos_pairs,
                    use_square, st);
}

int tdl_finish(float* S, float* Q, float* S_all, float* loss_b, float* coef_pos, float* coef_neg, float* tdl_out, int B,
               int Cs, int N, int D, float gamma_s, float gamma_d, int reverse_pos_pairs, int use_square, cudaStream_t st) {
  ProfScope prof(PT_TDL, st);
  TdlFlags f{gamma_s, gamma_d, reverse_pos_pairs, use_square};
  tdl_pair_kernel<<<B, 128, 0, st>>>(S, Q, S_all, loss_b, coef_pos, coef_neg, B, Cs, N, D, f);
  DCV_CUDA(cudaGetLastError());
  reduce_sum_kernel<<<1, 256, 0, st>>>(loss_b, B, 1.0f / static_cast<float>(B), tdl_out);
  DCV_CUDA(cudaGetLastError());
  count_launch(2);
  return 0;
}

int embed_bwd_dy(const float* G, const float* tokens, const float* addend, const float* bias, const float* rnorm,
                 const float* S, const float* S_all, const float* coef_pos, const float* coef_neg,
                 const float* d_extra, float lambda_tdl, void* dY, int B, int Cs, int N, int D, cudaStream_t st) {
  if (B <= 0 || Cs <= 0 || N <= 0) return set_error(DCV_ERR_INVALID, "embed_bwd_dy: empty problem");
  ProfScope prof(PT_EMBED_BWD, st);
  const long long rows = static_cast<long long>(B) * Cs * N;
  const int blocks = static_cast<int>(std::min<long long>((rows + 7) / 8, 148 * 16));
  DCV_NV_SWITCH(D, embed_bwd_dy_kernel<NV><<<blocks, 256, 0, st>>>(G, tokens, addend, bias, rnorm, S, S_all, coef_pos,
                                                                   coef_neg, d_extra, lambda_tdl,
                                                                   reinterpret_cast<__nv_bfloat16*>(dY), B, Cs, N, D));
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int embed_param_grads(const float* G, float* R, const int* gid, float* d_cls, float* d_pos0, float* d_chan_embed,
                      float* dpos_patch, int accumulate_pos, int B, int Cs, int N, int D, cudaStream_t st) {
  ProfScope prof(PT_EMBED_BWD, st);
  const long long L = static_cast<long long>(Cs) * N + 1;
  const long long LD4 = L * (D / 4);
  batch_sum_kernel<<<static_cast<unsigned>((LD4 + 255) / 256), 256, 0, st>>>(G, R, B, LD4);
  DCV_CUDA(cudaGetLastError());
  if (D % 4 || D > 1024) return set_error(DCV_ERR_UNSUPPORTED, "embed_param_grads: D=%d must be a multiple of 4, <= 1024", D);
  embed_param_grads_kernel<<<1 + Cs + N, D, static_cast<size_t>(D) * 4 * sizeof(float), st>>>(
      R, gid, d_cls, d_pos0, d_chan_embed, dpos_patch, accumulate_pos, Cs, N, D);
  DCV_CUDA(cudaGetLastError());
  count_launch(2);
  return 0;
}

int cdl_fwd(const float* chan_embed, const float* proxies, const int* gid, float scale, float* loss, float* dE,
            float* dP, int Cs, int D, cudaStream_t st) {
  if (Cs <= 0 || Cs > kCdlMaxC) return set_error(DCV_ERR_UNSUPPORTED, "cdl_fwd: C'=%d must be in [1,%d]", Cs, kCdlMaxC);
  ProfScope prof(PT_CDL, st);
  const size_t smem = static_cast<size_t>(2) * Cs * D * sizeof(float);
  // static + dynamic shared memory crosses the 48 KB default from C' = 14 (D = 384) on
  DCV_TRY_SMEM_ATTR(cdl_fwd_kernel, 200 * 1024);
  if (smem > 200 * 1024) return set_error(DCV_ERR_UNSUPPORTED, "cdl_fwd: C'*D too large for one CTA");
  cdl_fwd_kernel<<<1, 256, smem, st>>>(chan_embed, proxies, gid, scale, loss, dE, dP, Cs, D);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int cdl_bwd(const float* dE, const float* dP, const int* gid, const float* d_extra, float lambda_cdl,
            float* g_chan_embed, float* g_proxies, int Cs, int D, cudaStream_t st) {
  ProfScope prof(PT_CDL, st);
  cdl_bwd_kernel<<<(Cs * D + 255) / 256, 256, 0, st>>>(dE, dP, gid, d_extra, lambda_cdl, g_chan_embed, g_proxies, Cs, D);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int extra_loss(const float* tdl, const float* cdl, float lt, float lc, float* extra, cudaStream_t st) {
  ProfScope prof(PT_EMBED_MISC, st);
  extra_loss_kernel<<<1, 1, 0, st>>>(tdl, cdl, lt, lc, extra);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int sgemm_small(const float* A, int lda, int transA, const float* Bm, int ldb, int transB, float* C, int ldc,
                const float* bias, int accumulate, int M, int N, int K, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(DCV_ERR_INVALID, "sgemm_small: empty problem");
  ProfScope prof(PT_SMALL, st);
  dim3 grid((N + 31) / 32, (M + 31) / 32);
  sgemm_small_kernel<<<grid, 256, 0, st>>>(A, lda, transA, Bm, ldb, transB, C, ldc, bias, accumulate, M, N, K);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int cls_ln_fwd(const float* x, long long row_stride, const float* gamma, const float* beta, float* feat, float* mean,
               float* rstd, int B, int D, float eps, cudaStream_t st) {
  if (B <= 0) return set_error(DCV_ERR_INVALID, "cls_ln_fwd: empty");
  ProfScope prof(PT_SMALL, st);
  DCV_NV_SWITCH(D, cls_ln_fwd_kernel<NV><<<(B + 3) / 4, 128, 0, st>>>(x, row_stride, gamma, beta, feat, mean, rstd, B, D, eps));
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int cls_ln_bwd(const float* dfeat, const float* x, long long row_stride, const float* mean, const float* rstd,
               const float* gamma, float* dres, void* dres_bf16, float* dgamma, float* dbeta, float* dxsum, int B,
               int D, cudaStream_t st) {
  if (B <= 0) return set_error(DCV_ERR_INVALID, "cls_ln_bwd: empty");
  ProfScope prof(PT_SMALL, st);
  DCV_NV_SWITCH(D, cls_ln_bwd_kernel<NV><<<(B + 3) / 4, 128, 0, st>>>(dfeat, x, row_stride, mean, rstd, gamma, dres,
                                                                      reinterpret_cast<__nv_bfloat16*>(dres_bf16),
                                                                      dgamma, dbeta, dxsum, B, D));
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace dcv
