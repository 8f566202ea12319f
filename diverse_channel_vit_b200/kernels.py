"""Thin Python wrappers over the C ABI (include/dcvit.h): shape checks + pointer passing.

Every function enqueues work on the current CUDA stream and returns immediately.
PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESID, EPI_DGELU, EPI_F32 = 0, 1, 2, 3, 4


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.DcvError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.DcvError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise _lib.DcvError(f"{name}: innermost dimension must be contiguous")


def gemm_nt(a, b, epilogue=EPI_BIAS, bias=None, out=None, out2=None, resid=None, aux=None):
    """out = a[M,K] @ b[N,K]^T with fused epilogue (see dcvit.h DCV_EPI_*)."""
    _req(a, torch.bfloat16, "a"); _req(b, torch.bfloat16, "b")
    M, K = a.shape
    N, K2 = b.shape
    assert K == K2
    f32_out = epilogue in (EPI_BIAS_RESID, EPI_F32)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if f32_out else torch.bfloat16)
    if epilogue == EPI_BIAS_GELU and out2 is None:
        out2 = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    ldo = out.stride(0)
    for t in (out2, resid, aux):
        if t is not None:
            assert t.stride(0) == ldo
    check(_lib.lib().dcv_gemm_nt(ptr(a), a.stride(0), ptr(b), b.stride(0), M, N, K, epilogue, ptr(bias), ptr(out),
                                 ptr(out2), ptr(resid), ptr(aux), ldo, stream_ptr()), "dcv_gemm_nt")
    return (out, out2) if epilogue == EPI_BIAS_GELU else out


def gemm_nn(a, b, epilogue=EPI_BIAS, out=None, aux=None):
    """out = a[M,K] @ b[K,N] (b row-major, no transpose copy); epilogue BIAS(plain)/DGELU/F32."""
    _req(a, torch.bfloat16, "a"); _req(b, torch.bfloat16, "b")
    M, K = a.shape
    K2, N = b.shape
    assert K == K2
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if epilogue == EPI_F32 else torch.bfloat16)
    if aux is not None:
        assert aux.stride(0) == out.stride(0)
    check(_lib.lib().dcv_gemm_nn(ptr(a), a.stride(0), ptr(b), b.stride(0), M, N, K, epilogue, ptr(out), ptr(aux),
                                 out.stride(0), stream_ptr()), "dcv_gemm_nn")
    return out


def delta_ws(B, H, L, device, zero=False):
    """The `delta` workspace of the attention backward: fp32 [B, H, Lp]."""
    return (torch.zeros if zero else torch.empty)((B, H, lpad(L)), device=device, dtype=torch.float32)


def gemm_nn_delta(dy, w, o, B, L):
    """dO = dy[M,K] @ w[K,N] (bf16) and delta[B, N/64, Lp] = per-head rowsum(dO * o); M = B*L."""
    _req(dy, torch.bfloat16, "dy"); _req(w, torch.bfloat16, "w"); _req(o, torch.bfloat16, "o")
    M, K = dy.shape
    N = w.shape[1]
    assert M == B * L and o.shape == (M, N) and o.is_contiguous()
    d_o = torch.empty((M, N), device=dy.device, dtype=torch.bfloat16)
    delta = delta_ws(B, N // 64, L, dy.device, zero=True)
    check(_lib.lib().dcv_gemm_nn_delta(ptr(dy), dy.stride(0), ptr(w), w.stride(0), M, N, K, ptr(d_o), ptr(o), ptr(delta),
                                       L, stream_ptr()), "dcv_gemm_nn_delta")
    return d_o, delta


def gemm_tn(a, b, out=None, accumulate=True, splits=0):
    """out[Nout,Kout] (+)= a[M,Nout]^T @ b[M,Kout]; fp32 output."""
    _req(a, torch.bfloat16, "a"); _req(b, torch.bfloat16, "b")
    M, Nout = a.shape
    M2, Kout = b.shape
    assert M == M2
    if out is None:
        out = torch.zeros((Nout, Kout), device=a.device, dtype=torch.float32)
    check(_lib.lib().dcv_gemm_tn(ptr(a), a.stride(0), ptr(b), b.stride(0), M, Nout, Kout, ptr(out), out.stride(0),
                                 1 if accumulate else 0, splits, stream_ptr()), "dcv_gemm_tn")
    return out


def lpad(L: int) -> int:
    return (L + 127) // 128 * 128


def attn_fwd(qkv, B, L, H, scale=None, o=None, lse2=None):
    """qkv bf16 [B*L, 3*H*64] -> (o bf16 [B*L, H*64], lse2 fp32 [B,H,L])."""
    _req(qkv, torch.bfloat16, "qkv")
    D = H * 64
    assert qkv.is_contiguous() and qkv.numel() == B * L * 3 * D
    if o is None:
        o = torch.empty((B * L, D), device=qkv.device, dtype=torch.bfloat16)
    if lse2 is None:
        lse2 = torch.empty((B, H, lpad(L)), device=qkv.device, dtype=torch.float32)
    scale = 64 ** -0.5 if scale is None else scale
    check(_lib.lib().dcv_attn_fwd(ptr(qkv), ptr(o), ptr(lse2), B, L, H, ctypes.c_float(scale), stream_ptr()),
          "dcv_attn_fwd")
    return o, lse2


def attn_bwd(qkv, o, do, lse2, B, L, H, scale=None, dqkv=None, delta=None, dq_acc=None, dbias=None, delta_ready=False):
    """-> dqkv bf16 [B*L, 3*H*64].  dbias (fp32 [3*H*64]): the column sums of dqkv are added to it by the kernels;
    delta_ready: `delta` already holds rowsum(dO*O) per head (gemm_nn_delta)."""
    for t, n in ((qkv, "qkv"), (o, "o"), (do, "do")):
        _req(t, torch.bfloat16, n)
        assert t.is_contiguous()
    D = H * 64
    dev = qkv.device
    if dqkv is None:
        dqkv = torch.empty((B * L, 3 * D), device=dev, dtype=torch.bfloat16)
    if delta is None:
        assert not delta_ready
        delta = delta_ws(B, H, L, dev)
    if dq_acc is None:
        dq_acc = torch.empty((B, H, L, 64), device=dev, dtype=torch.float32)
    scale = 64 ** -0.5 if scale is None else scale
    if dbias is None and not delta_ready:
        check(_lib.lib().dcv_attn_bwd(ptr(qkv), ptr(o), ptr(do), ptr(lse2), ptr(delta), ptr(dq_acc), ptr(dqkv), B, L, H,
                                      ctypes.c_float(scale), stream_ptr()), "dcv_attn_bwd")
    else:
        check(_lib.lib().dcv_attn_bwd_fused(ptr(qkv), ptr(o), ptr(do), ptr(lse2), ptr(delta), ptr(dq_acc), ptr(dqkv),
                                            ptr(dbias), int(delta_ready), B, L, H, ctypes.c_float(scale), stream_ptr()),
              "dcv_attn_bwd_fused")
    return dqkv

def ln_fwd(x, gamma, beta, eps=1e-6):
    """x fp32 [M,D] -> (y bf16, mean, rstd)."""
    _req(x, torch.float32, "x")
    M, D = x.shape
    y = torch.empty((M, D), device=x.device, dtype=torch.bfloat16)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    check(_lib.lib().dcv_ln_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), M, D, ctypes.c_float(eps),
                                stream_ptr()), "dcv_ln_fwd")
    return y, mean, rstd


def ln_bwd(dy, x, mean, rstd, gamma, dres, dgamma, dbeta, dxsum=None):
    """dres (fp32, in place) += LN'(dy); returns the bf16 copy of the updated dres."""
    _req(dy, torch.bfloat16, "dy"); _req(x, torch.float32, "x"); _req(dres, torch.float32, "dres")
    M, D = x.shape
    dxb = torch.empty((M, D), device=x.device, dtype=torch.bfloat16)
    check(_lib.lib().dcv_ln_bwd(ptr(dy), ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(dres), ptr(dxb), ptr(dgamma),
                                ptr(dbeta), ptr(dxsum), M, D, stream_ptr()), "dcv_ln_bwd")
    return dxb


def colsum_bf16(a, out):
    _req(a, torch.bfloat16, "a")
    check(_lib.lib().dcv_colsum_bf16(ptr(a), ptr(out), a.shape[0], a.shape[1], a.stride(0), stream_ptr()), "dcv_colsum_bf16")
    return out


def cast_f32_bf16(src):
    _req(src, torch.float32, "src")
    dst = torch.empty_like(src, dtype=torch.bfloat16)
    check(_lib.lib().dcv_cast_f32_bf16(ptr(src), ptr(dst), ctypes.c_longlong(src.numel()), stream_ptr()), "dcv_cast_f32_bf16")
    return dst


def sgemm_small(a, b, bias=None, out=None, accumulate=False, trans_a=False, trans_b=False):
    """fp32 C = op(a) op(b) (+bias); trans_a: a stored [K,M]; trans_b: b stored [N,K]."""
    _req(a, torch.float32, "a"); _req(b, torch.float32, "b")
    M, Kd = (a.shape[1], a.shape[0]) if trans_a else a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    if out is None:
        out = torch.zeros((M, N), device=a.device, dtype=torch.float32)
    check(_lib.lib().dcv_sgemm_small(ptr(a), a.stride(0), int(trans_a), ptr(b), b.stride(0), int(trans_b), ptr(out),
                                     out.stride(0), ptr(bias), int(accumulate), M, N, Kd, stream_ptr()), "dcv_sgemm_small")
    return out
