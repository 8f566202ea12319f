// HBM-bound row kernels of the transformer blocks: LayerNorm forward/backward
// (reference models/vit.py:384,398 norm1/norm2 and models/dichavit.py:651 norm;
// nn.LayerNorm, eps 1e-6, biased variance), bias-gradient column sums, and the
// fp32 -> bf16 parameter cast.  One warp owns one row; all loads/stores are 8- or
// 16-byte vectors, coalesced across the warp.
#include <algorithm>

#include "common.cuh"
#include "host.h"

namespace dcv {

constexpr int kRowWarps = 8;  // warps per CTA

// ---------------------------------------------------------------------------------
// LayerNorm forward: x fp32 [M,D] -> y bf16 [M,D], mean/rstd fp32 [M]
// ---------------------------------------------------------------------------------
template <int NV>  // float4 per lane, NV = ceil(D / 128)
__global__ void __launch_bounds__(kRowWarps * 32, NV <= 3 ? 4 : 2)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int M, int D,
              float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 2;
  float4 g[NV], bt[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int idx = lane + 32 * k;
    if (idx < nvec) {
      g[k] = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
      bt[k] = __ldg(reinterpret_cast<const float4*>(beta) + idx);
    } else {
      g[k] = make_float4(0, 0, 0, 0);
      bt[k] = make_float4(0, 0, 0, 0);
    }
  }
  const float inv_d = 1.0f / static_cast<float>(D);
  // Software pipeline: the next row's loads are in flight while this row goes through its two dependent warp
  // reductions and its stores (the un-pipelined loop left a warp with nothing outstanding for ~2/3 of a row's time:
  // 0.71 of the HBM copy bandwidth against 0.87 for the pipelined backward kernel).
  const int stride = gridDim.x * kRowWarps;
  int row = blockIdx.x * kRowWarps + warp;
  float4 nx[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int idx = lane + 32 * k;
    nx[k] = (row < M && idx < nvec) ? reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D)[idx]
                                    : make_float4(0, 0, 0, 0);
  }
  for (; row < M; row += stride) {
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      v[k] = nx[k];
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const int next = row + stride;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int idx = lane + 32 * k;
      if (next < M && idx < nvec) nx[k] = reinterpret_cast<const float4*>(x + static_cast<size_t>(next) * D)[idx];
    }
    const float mu = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int idx = lane + 32 * k;
      if (idx < nvec) {
        const float a = v[k].x - mu, b = v[k].y - mu, c = v[k].z - mu, d = v[k].w - mu;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rs = rsqrtf(warp_sum(q) * inv_d + eps);
    uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * D);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int idx = lane + 32 * k;
      if (idx < nvec) {
        uint2 o;
        o.x = pack_bf16((v[k].x - mu) * rs * g[k].x + bt[k].x, (v[k].y - mu) * rs * g[k].y + bt[k].y);
        o.y = pack_bf16((v[k].z - mu) * rs * g[k].z + bt[k].z, (v[k].w - mu) * rs * g[k].w + bt[k].w);
        yr[idx] = o;
      }
    }
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

// ---------------------------------------------------------------------------------
// LayerNorm backward, fused with the residual-stream gradient add:
//   dx_out = dres + LN'(dy)       (fp32, in place on dres) and a bf16 copy for the next GEMMs
//   dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy ; dxsum += sum_rows dx_out (bias grad of the
//   Linear whose output fed this residual add)
// Persistent grid; per-lane column accumulators, one smem reduction + D atomics per CTA.
// ---------------------------------------------------------------------------------
template <int NV>
struct LnBwdRow {  // one row's operands as they come from memory
  float4 x[NV], dr[NV];
  uint2 dy[NV];
  float mu, rs;
};

template <int NV>
__device__ __forceinline__ void ln_bwd_load(LnBwdRow<NV>& t, const float* __restrict__ x,
                                            const __nv_bfloat16* __restrict__ dy, const float* __restrict__ dres,
                                            const float* __restrict__ mean, const float* __restrict__ rstd, int row,
                                            int D, int nvec, int lane) {
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
  const uint2* dyr = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(row) * D);
  const float4* dr = reinterpret_cast<const float4*>(dres + static_cast<size_t>(row) * D);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int idx = lane + 32 * k;
    if (idx < nvec) {
      t.x[k] = __ldcs(xr + idx);   // streamed once: do not keep in L1/L2 longer than needed
      t.dy[k] = __ldcs(dyr + idx);
      t.dr[k] = __ldcs(dr + idx);
    } else {
      t.x[k] = t.dr[k] = make_float4(0, 0, 0, 0);
      t.dy[k] = make_uint2(0, 0);
    }
  }
  t.mu = mean[row];
  t.rs = rstd[row];
}

template <int NV>
__global__ void __launch_bounds__(kRowWarps * 32, NV <= 3 ? 2 : 1)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, float* __restrict__ dres,
              __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dxsum, int M, int D) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float red[];  // [kRowWarps][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 2;
  float4 g[NV], acc_g[NV], acc_b[NV], acc_s[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int idx = lane + 32 * k;
    g[k] = idx < nvec ? __ldg(reinterpret_cast<const float4*>(gamma) + idx) : make_float4(0, 0, 0, 0);
    acc_g[k] = acc_b[k] = acc_s[k] = make_float4(0, 0, 0, 0);
  }
  const float inv_d = 1.0f / static_cast<float>(D);
  const int stride = gridDim.x * kRowWarps;
  int row = blockIdx.x * kRowWarps + warp;
  LnBwdRow<NV> cur, nxt;
  if (row < M) ln_bwd_load(cur, x, dy, dres, mean, rstd, row, D, nvec, lane);
  for (; row < M; row += stride) {
    // software pipeline: the next row's loads are in flight while this row is reduced and stored
    if (row + stride < M) ln_bwd_load(nxt, x, dy, dres, mean, rstd, row + stride, D, nvec, lane);
    float4 xh[NV], gy[NV], dyv[NV];
    float c1 = 0.f, c2 = 0.f;
    const float mu = cur.mu, rs = cur.rs;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const float2 d01 = unpack_bf16(cur.dy[k].x), d23 = unpack_bf16(cur.dy[k].y);
      dyv[k] = make_float4(d01.x, d01.y, d23.x, d23.y);
      xh[k] = make_float4((cur.x[k].x - mu) * rs, (cur.x[k].y - mu) * rs, (cur.x[k].z - mu) * rs, (cur.x[k].w - mu) * rs);
      gy[k] = make_float4(dyv[k].x * g[k].x, dyv[k].y * g[k].y, dyv[k].z * g[k].z, dyv[k].w * g[k].w);
      if (lane + 32 * k < nvec) {
        c1 += (gy[k].x + gy[k].y) + (gy[k].z + gy[k].w);
        c2 += (gy[k].x * xh[k].x + gy[k].y * xh[k].y) + (gy[k].z * xh[k].z + gy[k].w * xh[k].w);
      }
    }
    c1 = warp_sum(c1) * inv_d;
    c2 = warp_sum(c2) * inv_d;
    float4* dr = reinterpret_cast<float4*>(dres + static_cast<size_t>(row) * D);
    uint2* dxb = reinterpret_cast<uint2*>(dx_bf16 + static_cast<size_t>(row) * D);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int idx = lane + 32 * k;
      if (idx < nvec) {
        float4 o = cur.dr[k];
        o.x += rs * (gy[k].x - c1 - xh[k].x * c2);
        o.y += rs * (gy[k].y - c1 - xh[k].y * c2);
        o.z += rs * (gy[k].z - c1 - xh[k].z * c2);
        o.w += rs * (gy[k].w - c1 - xh[k].w * c2);
        dr[idx] = o;
        uint2 ob;
        ob.x = pack_bf16(o.x, o.y);
        ob.y = pack_bf16(o.z, o.w);
        dxb[idx] = ob;
        acc_g[k].x += dyv[k].x * xh[k].x; acc_g[k].y += dyv[k].y * xh[k].y;
        acc_g[k].z += dyv[k].z * xh[k].z; acc_g[k].w += dyv[k].w * xh[k].w;
        acc_b[k].x += dyv[k].x; acc_b[k].y += dyv[k].y; acc_b[k].z += dyv[k].z; acc_b[k].w += dyv[k].w;
        acc_s[k].x += o.x; acc_s[k].y += o.y; acc_s[k].z += o.z; acc_s[k].w += o.w;
      }
    }
    cur = nxt;
  }
  // CTA reduction of the three column accumulators, one after the other through `red`
#pragma unroll 1
  for (int which = 0; which < 3; ++which) {
    float* target = which == 0 ? dgamma : (which == 1 ? dbeta : dxsum);
    if (target == nullptr) continue;  // uniform across the CTA
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int idx = lane + 32 * k;
      if (idx < nvec) {
        const float4 a = which == 0 ? acc_g[k] : (which == 1 ? acc_b[k] : acc_s[k]);
        reinterpret_cast<float4*>(red + warp * D)[idx] = a;
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) t += red[w * D + c];
      atomicAdd(target + c, t);
    }
  }
}

// ---------------------------------------------------------------------------------
// column sums of a bf16 matrix: out[n] += sum_m a[m,n]   (bias gradients of qkv / fc1)
// grid (ceil(N/256), row_splits); each warp reads 512 contiguous bytes of a row
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowWarps * 32)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ a, float* __restrict__ out, int M, int N, int lda) {
  __shared__ float red[kRowWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 256 + lane * 8;
  const bool col_ok = col < N;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per;
  const int r1 = min(M, r0 + rows_per);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col_ok) {
    for (int r = r0 + warp; r < r1; r += kRowWarps) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(a + static_cast<size_t>(r) * lda + col));
      const float2 a0 = unpack_bf16(v.x), a1 = unpack_bf16(v.y), a2 = unpack_bf16(v.z), a3 = unpack_bf16(v.w);
      acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y;
      acc[4] += a2.x; acc[5] += a2.y; acc[6] += a3.x; acc[7] += a3.y;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;  // 256 threads == 256 columns
  if (blockIdx.x * 256 + c < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) t += red[w][c];
    atomicAdd(out + blockIdx.x * 256 + c, t);
  }
}

// out[n] += sum_m a[m,n], a fp32 [M,N] with a handful of rows (classifier-head bias gradient)
__global__ void colsum_f32_kernel(const float* __restrict__ a, float* __restrict__ out, int M, int N, int lda) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float t = 0.f;
  for (int m = 0; m < M; ++m) t += a[static_cast<size_t>(m) * lda + n];
  out[n] += t;
}

// ---------------------------------------------------------------------------------
// fp32 -> bf16 cast of the flat parameter buffer (once per optimiser step)
// ---------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + i) = o;
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
  }
}

// ---------------------------------------------------------------------------------
// Fused AdamW over the flat parameter / gradient buffers (reference optimizers.py:20-21 timm AdamW ==
// torch.optim.AdamW semantics: decoupled weight decay, bias-corrected moments), one launch per step, 16 B/param
// read + 12 B/param written (+2 for the bf16 operand copy of the next forward).  `clip`: optional device
// scalar pair {sum of squared gradients, max_norm}: gradients are scaled by min(1, max_norm / (norm + 1e-6))
// like torch.nn.utils.clip_grad_norm_ (trainer.py:1003-1004).
// ---------------------------------------------------------------------------------
struct AdamWArgs {
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt;
};

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ p_bf16, long long n4, AdamWArgs a, const float* __restrict__ clip) {
  float gs = 1.0f;
  if (clip != nullptr) {
    const float norm = sqrtf(clip[0]);
    gs = fminf(1.0f, clip[1] / (norm + 1e-6f));
  }
  const float step_size = a.lr / a.bc1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gp = reinterpret_cast<const float*>(&gv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = gp[k] * gs;
      pp[k] *= 1.0f - a.lr * a.wd;
      mp[k] = a.beta1 * mp[k] + (1.0f - a.beta1) * gr;
      vp[k] = a.beta2 * vp[k] + (1.0f - a.beta2) * gr * gr;
      const float denom = sqrtf(vp[k]) / a.bc2_sqrt + a.eps;
      pp[k] -= step_size * (mp[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (p_bf16 != nullptr) {
      uint2 o;
      o.x = pack_bf16(pp[0], pp[1]);
      o.y = pack_bf16(pp[2], pp[3]);
      reinterpret_cast<uint2*>(p_bf16)[i] = o;
    }
  }
}

// Same update with every per-step scalar read from DEVICE memory (dcv_optim_state), so that a captured CUDA graph of
// the optimiser step replays with the learning rate / weight decay / bias corrections of the current update.
__global__ void __launch_bounds__(256)
adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 __nv_bfloat16* __restrict__ p_bf16, long long n4, float beta1, float beta2, float eps,
                 const dcv_optim_state* __restrict__ state, const float* __restrict__ clip) {
  float gs = 1.0f;
  if (clip != nullptr) {
    const float norm = sqrtf(clip[0]);
    gs = fminf(1.0f, clip[1] / (norm + 1e-6f));
  }
  const float lr = state->lr, wd = state->wd, bc2_sqrt = state->bc2_sqrt;
  const float step_size = lr / state->bc1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gp = reinterpret_cast<const float*>(&gv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = gp[k] * gs;
      pp[k] *= 1.0f - lr * wd;
      mp[k] = beta1 * mp[k] + (1.0f - beta1) * gr;
      vp[k] = beta2 * vp[k] + (1.0f - beta2) * gr * gr;
      const float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
      pp[k] -= step_size * (mp[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (p_bf16 != nullptr) {
      uint2 o;
      o.x = pack_bf16(pp[0], pp[1]);
      o.y = pack_bf16(pp[2], pp[3]);
      reinterpret_cast<uint2*>(p_bf16)[i] = o;
    }
  }
}

// One thread: advance the update counter and evaluate the schedules for the update about to be applied.
//   lr  : timm CosineLRScheduler._get_lr(t) (reference lr_schedulers.py:6-9), t = u - 1 when the schedule counts
//         updates (trainer.py:1009-1010 sets the next update's lr right after optimizer.step()) or the 1-based epoch
//         number 1 + (u - 1) / updates_per_epoch when it counts epochs (trainer.py:344-348)
//   wd  : utils.cosine_scheduler table (utils.py:563-574) indexed as trainer.py:1011-1019 does: update 1 runs with the
//         optimiser's own weight decay (= table[0]), update u >= 2 with table[min(u - 2, len - 1)]
__global__ void optim_sched_kernel(dcv_optim_state* st, dcv_sched c) {
  const int u = st->num_updates + 1;
  st->num_updates = u;
  const int t = c.updates_per_epoch > 0 ? 1 + (u - 1) / c.updates_per_epoch : u - 1;
  float lr = c.base_lr;
  if (c.t_initial > 0) {
    if (t < c.warmup_t) {
      lr = c.warmup_lr_init + t * ((c.base_lr - c.warmup_lr_init) / c.warmup_t);
    } else {
      const int tt = c.warmup_prefix ? t - c.warmup_t : t;
      const int i = tt / c.t_initial;
      const int t_curr = tt - c.t_initial * i;
      const double lr_max = static_cast<double>(c.base_lr) * pow(static_cast<double>(c.cycle_decay), static_cast<double>(i));
      if (i < c.cycle_limit)
        lr = static_cast<float>(c.lr_min + 0.5 * (lr_max - c.lr_min) *
                                (1.0 + cos(3.14159265358979323846 * pow(static_cast<double>(t_curr), static_cast<double>(c.k_decay)) /
                                           pow(static_cast<double>(c.t_initial), static_cast<double>(c.k_decay)))));
      else
        lr = c.lr_min;
    }
  }
  float wd = c.wd_base;
  if (c.wd_total > 0 && u >= 2) {
    const int idx = min(u - 2, c.wd_total - 1);
    wd = static_cast<float>(c.wd_end + 0.5 * (static_cast<double>(c.wd_base) - c.wd_end) *
                            (1.0 + cos(3.14159265358979323846 * idx / static_cast<double>(c.wd_total))));
  }
  st->lr = lr;
  st->wd = wd;
  st->bc1 = static_cast<float>(1.0 - pow(static_cast<double>(c.beta1), static_cast<double>(u)));
  st->bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(c.beta2), static_cast<double>(u))));
}

// out[0] += sum(g^2)   (global gradient norm for clipping; out zeroed by the caller)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n4, float* __restrict__ out) {
  __shared__ float red[8];
  float t = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    t += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(out, s);
  }
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
static int nv_for(int D) { return (D + 127) / 128; }

int ln_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M, int D,
           float eps, cudaStream_t st) {
  if (M <= 0 || D <= 0) return set_error(DCV_ERR_INVALID, "ln_fwd: empty problem");
  if (D % 4) return set_error(DCV_ERR_UNSUPPORTED, "ln_fwd: D=%d must be a multiple of 4", D);
  ProfScope prof(PT_LN_FWD, st);
  // one resident wave: 4 CTAs of 8 warps per SM (<= 64 registers; 2 CTAs for D = 768, whose rows need twice the
  // registers), every warp walks its rows with a one-row prefetch
  const int blocks = min((M + kRowWarps - 1) / kRowWarps, num_sms() * (nv_for(D) <= 3 ? 4 : 2));
  __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(y);
  switch (nv_for(D)) {
    case 1: DCV_CUDA(launch_pdl(ln_fwd_kernel<1>, dim3(blocks), dim3(kRowWarps * 32), 0, st, x, gamma, beta, yb, mean, rstd, M, D, eps)); break;
    case 2: DCV_CUDA(launch_pdl(ln_fwd_kernel<2>, dim3(blocks), dim3(kRowWarps * 32), 0, st, x, gamma, beta, yb, mean, rstd, M, D, eps)); break;
    case 3: DCV_CUDA(launch_pdl(ln_fwd_kernel<3>, dim3(blocks), dim3(kRowWarps * 32), 0, st, x, gamma, beta, yb, mean, rstd, M, D, eps)); break;
    case 6: DCV_CUDA(launch_pdl(ln_fwd_kernel<6>, dim3(blocks), dim3(kRowWarps * 32), 0, st, x, gamma, beta, yb, mean, rstd, M, D, eps)); break;
    default: return set_error(DCV_ERR_UNSUPPORTED, "ln_fwd: D=%d not instantiated (128/256/384/768 classes)", D);
  }
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int ln_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* gamma, float* dres,
           void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, int M, int D, cudaStream_t st) {
  if (M <= 0 || D <= 0) return set_error(DCV_ERR_INVALID, "ln_bwd: empty problem");
  if (D % 4) return set_error(DCV_ERR_UNSUPPORTED, "ln_bwd: D=%d must be a multiple of 4", D);
  ProfScope prof(PT_LN_BWD, st);
  const int blocks = min((M + kRowWarps - 1) / kRowWarps, num_sms() * 4);
  const size_t smem = static_cast<size_t>(kRowWarps) * D * sizeof(float);
  const __nv_bfloat16* dyb = reinterpret_cast<const __nv_bfloat16*>(dy);
  __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
#define LNB(NV)                                                                                                  \
  DCV_CUDA(launch_pdl(ln_bwd_kernel<NV>, dim3(blocks), dim3(kRowWarps * 32), smem, st, dyb, x, mean, rstd, gamma, dres, \
                      dxb, dgamma, dbeta, dxsum, M, D))
  switch (nv_for(D)) {
    case 1: LNB(1); break;
    case 2: LNB(2); break;
    case 3: LNB(3); break;
    case 6: LNB(6); break;
    default: return set_error(DCV_ERR_UNSUPPORTED, "ln_bwd: D=%d not instantiated", D);
  }
#undef LNB
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int colsum_bf16(const void* a, float* out, int M, int N, int lda, cudaStream_t st) {
  if (M <= 0 || N <= 0) return set_error(DCV_ERR_INVALID, "colsum: empty problem");
  if (N % 8 || lda % 8) return set_error(DCV_ERR_UNSUPPORTED, "colsum: N and lda must be multiples of 8");
  ProfScope prof(PT_COLSUM, st);
  const int cb = (N + 255) / 256;
  int splits = (num_sms() * 4 + cb - 1) / cb;
  if (splits > (M + 31) / 32) splits = (M + 31) / 32;
  if (splits < 1) splits = 1;
  colsum_bf16_kernel<<<dim3(cb, splits), kRowWarps * 32, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(a), out, M,
                                                                  N, lda);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int colsum_f32(const float* a, float* out, int M, int N, int lda, cudaStream_t st) {
  if (M <= 0 || N <= 0) return set_error(DCV_ERR_INVALID, "colsum_f32: empty problem");
  ProfScope prof(PT_SMALL, st);
  colsum_f32_kernel<<<(N + 127) / 128, 128, 0, st>>>(a, out, M, N, lda);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
               float beta2, float eps, float wd, int step, const float* clip, cudaStream_t st) {
  if (n <= 0 || (n & 3)) return set_error(DCV_ERR_INVALID, "adamw_step: n must be a positive multiple of 4");
  if (step < 1) return set_error(DCV_ERR_INVALID, "adamw_step: step counts from 1");
  ProfScope prof(PT_CAST, st);
  AdamWArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = wd;
  a.bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  a.bc2_sqrt = sqrtf(1.0f - powf(beta2, static_cast<float>(step)));
  const long long n4 = n >> 2;
  const int blocks = static_cast<int>(std::min<long long>((n4 + 255) / 256, static_cast<long long>(num_sms()) * 16));
  adamw_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16), n4, a, clip);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int adamw_step_dev(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float beta1, float beta2,
                   float eps, const dcv_optim_state* state, const float* clip, cudaStream_t st) {
  if (n <= 0 || (n & 3)) return set_error(DCV_ERR_INVALID, "adamw_step_dev: n must be a positive multiple of 4");
  ProfScope prof(PT_CAST, st);
  const long long n4 = n >> 2;
  const int blocks = static_cast<int>(std::min<long long>((n4 + 255) / 256, static_cast<long long>(num_sms()) * 16));
  adamw_dev_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16), n4, beta1, beta2, eps,
                                           state, clip);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int optim_sched_step(dcv_optim_state* state, const dcv_sched& cfg, cudaStream_t st) {
  if (cfg.t_initial > 0 && cfg.warmup_t < 0) return set_error(DCV_ERR_INVALID, "optim_sched_step: warmup_t < 0");
  optim_sched_kernel<<<1, 1, 0, st>>>(state, cfg);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int sumsq_f32(const float* g, long long n, float* out, cudaStream_t st) {
  if (n <= 0 || (n & 3)) return set_error(DCV_ERR_INVALID, "sumsq: n must be a positive multiple of 4");
  ProfScope prof(PT_CAST, st);
  const long long n4 = n >> 2;
  const int blocks = static_cast<int>(std::min<long long>((n4 + 255) / 256, static_cast<long long>(num_sms()) * 8));
  sumsq_kernel<<<blocks, 256, 0, st>>>(g, n4, out);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int cast_f32_bf16(const float* src, void* dst, long long n, cudaStream_t st) {
  if (n <= 0) return set_error(DCV_ERR_INVALID, "cast: empty");
  ProfScope prof(PT_CAST, st);
  const long long threads = (n + 3) / 4;
  cast_f32_bf16_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  DCV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace dcv
