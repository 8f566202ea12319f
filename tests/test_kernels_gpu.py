"""GPU: every primitive of the C ABI against a plain fp32 torch reference of the same op.
Tolerances: bf16 operands / bf16 outputs -> rel-L2 <= 1e-2 (north_star activation tolerance);
fp32-accumulated outputs (wgrad, residual epilogue, LayerNorm statistics) much tighter."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from diverse_channel_vit_b200 import kernels

    return kernels


def _bf(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).bfloat16()


@pytest.mark.parametrize("M,N,Kd", [(128, 192, 64), (1000, 1152, 384), (777, 128, 1536), (300, 64, 128), (50, 576, 192)])
def test_gemm_nt_bias(K, M, N, Kd):
    a, b = _bf(M, Kd, seed=1), _bf(N, Kd, scale=0.05, seed=2)
    bias = torch.randn(N, device="cuda")
    out = K.gemm_nt(a, b, K.EPI_BIAS, bias=bias)
    assert rel_l2(out, a.float() @ b.float().t() + bias) < 1e-2


def test_gemm_nt_epilogues(K):
    M, N, Kd = 520, 384, 384
    a, b = _bf(M, Kd, seed=3), _bf(N, Kd, scale=0.05, seed=4)
    bias = torch.randn(N, device="cuda")
    ref = a.float() @ b.float().t() + bias
    h, g = K.gemm_nt(a, b, K.EPI_BIAS_GELU, bias=bias)
    assert rel_l2(h, ref) < 1e-2 and rel_l2(g, F.gelu(ref)) < 1e-2
    res = torch.randn(M, N, device="cuda")
    out = K.gemm_nt(a, b, K.EPI_BIAS_RESID, bias=bias, resid=res)
    assert rel_l2(out, ref + res) < 1e-5
    hh = _bf(M, N, seed=5)
    dg = K.gemm_nt(a, b, K.EPI_DGELU, aux=hh)
    hf = hh.float().requires_grad_(True)
    F.gelu(hf).backward(a.float() @ b.float().t())
    assert rel_l2(dg, hf.grad) < 1e-2


@pytest.mark.parametrize("M,N,Kd", [(333, 384, 1536), (1000, 1536, 384), (200, 192, 576)])
def test_gemm_nn_dgrad(K, M, N, Kd):
    a, w = _bf(M, Kd, seed=6), _bf(Kd, N, scale=0.05, seed=7)
    out = K.gemm_nn(a, w)
    assert rel_l2(out, a.float() @ w.float()) < 1e-2
    hh = _bf(M, N, seed=8)
    dg = K.gemm_nn(a, w, K.EPI_DGELU, aux=hh)
    hf = hh.float().requires_grad_(True)
    F.gelu(hf).backward(a.float() @ w.float())
    assert rel_l2(dg, hf.grad) < 1e-2


@pytest.mark.parametrize("B,L,H", [(2, 197, 3), (3, 81, 6), (1, 1569, 6), (4, 33, 2)])
def test_attn_bwd_fused_bias_gradient_and_delta(K, B, L, H):
    """the block backward's fused path: delta from the projection-dgrad epilogue, qkv bias gradient from the
    attention-backward epilogues == column sums of the dqkv the plain path stores"""
    D = H * 64
    qkv, dy, w = _bf(B * L, 3 * D, seed=21), _bf(B * L, D, seed=22), _bf(D, D, scale=0.05, seed=23)
    o, lse = K.attn_fwd(qkv, B, L, H)
    d_o, delta = K.gemm_nn_delta(dy, w, o, B, L)
    ref = K.attn_bwd(qkv, o, d_o, lse, B, L, H)
    dbias = torch.full((3 * D,), 0.25, device="cuda")  # accumulates on top of what is there
    got = K.attn_bwd(qkv, o, d_o, lse, B, L, H, delta=delta, delta_ready=True, dbias=dbias)
    assert rel_l2(got, ref) < 2e-3  # same kernels; dQ accumulation order (fp32 reductions) differs run to run
    assert rel_l2(dbias - 0.25, got.float().sum(0)) < 5e-3  # fp32 sums before the bf16 rounding vs sums of rounded values


@pytest.mark.parametrize("B,L,H", [(2, 197, 6), (3, 81, 3), (1, 1569, 6)])
def test_gemm_nn_delta(K, B, L, H):
    """projection dgrad with the fused softmax-backward row term delta = rowsum_head(dO * O)"""
    D = H * 64
    dy, w, o = _bf(B * L, D, seed=11), _bf(D, D, scale=0.05, seed=12), _bf(B * L, D, seed=13)
    d_o, delta = K.gemm_nn_delta(dy, w, o, B, L)
    ref = dy.float() @ w.float()
    assert rel_l2(d_o, ref) < 1e-2
    dref = (d_o.float() * o.float()).reshape(B, L, H, 64).sum(-1).permute(0, 2, 1)  # from the rounded dO, as the kernel
    assert rel_l2(delta[:, :, :L], dref) < 1e-5
    assert delta[:, :, L:].abs().max().item() == 0.0


@pytest.mark.parametrize("M,Nout,Kout", [(64, 128, 192), (5000, 1152, 384), (3001, 384, 1536), (515, 64, 64), (999, 384, 256)])
def test_gemm_tn_wgrad(K, M, Nout, Kout):
    a, b = _bf(M, Nout, seed=9), _bf(M, Kout, seed=10)
    out = K.gemm_tn(a, b)
    assert rel_l2(out, a.float().t() @ b.float()) < 1e-5
    # accumulation into a running gradient
    K.gemm_tn(a, b, out=out)
    assert rel_l2(out, 2 * (a.float().t() @ b.float())) < 1e-5


def _attn_ref(qkv, B, L, H):
    D = H * 64
    q, k, v = qkv.float().reshape(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-2, -1)) * 0.125
    p = s.softmax(-1)
    return (p @ v).transpose(1, 2).reshape(B * L, D), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,L,H", [(1, 128, 1), (2, 197, 3), (2, 589, 6), (1, 1569, 6), (3, 81, 3), (2, 17, 3), (2, 1, 2)])
def test_attention_fwd_bwd(K, B, L, H):
    D = H * 64
    qkv = _bf(B * L, 3 * D, seed=11)
    o, lse2 = K.attn_fwd(qkv, B, L, H)
    qf = qkv.float().requires_grad_(True)
    ref, lse = _attn_ref(qf, B, L, H)
    assert rel_l2(o, ref) < 1e-2
    assert rel_l2(lse2[:, :, :L] * 0.6931471805599453, lse) < 1e-5
    do = _bf(B * L, D, seed=12)
    ref.backward(do.float())
    dqkv = K.attn_bwd(qkv, o, do, lse2, B, L, H)
    for i, nm in enumerate("qkv"):
        got = dqkv.float().reshape(B, L, 3, D)[:, :, i]
        want = qf.grad.reshape(B, L, 3, D)[:, :, i]
        if want.abs().max() < 1e-12:  # L == 1: softmax of one element, dq = dk = 0 exactly in the reference
            assert got.abs().max().item() < 1e-6, nm
        else:
            assert rel_l2(got, want) < 1.5e-2, nm


@pytest.mark.parametrize("M,D", [(1000, 384), (77, 192), (513, 768)])
def test_layernorm_fwd_bwd(K, M, D):
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(M, D, device="cuda", generator=g) * 2 + 0.3
    gamma = torch.randn(D, device="cuda", generator=g) * 0.1 + 1
    beta = torch.randn(D, device="cuda", generator=g) * 0.1
    y, mean, rstd = K.ln_fwd(x, gamma, beta)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (D,), gr, br, eps=1e-6)
    assert rel_l2(y, ref) < 5e-3
    assert rel_l2(mean, x.mean(1)) < 1e-5 and rel_l2(rstd, (x.var(1, unbiased=False) + 1e-6).rsqrt()) < 1e-5
    dy = _bf(M, D, seed=14)
    dres = torch.randn(M, D, device="cuda", generator=g)
    ref.backward(dy.float())
    want_dx = dres + xr.grad
    dgam, dbet, dsum = (torch.zeros(D, device="cuda") for _ in range(3))
    dxb = K.ln_bwd(dy, x, mean, rstd, gamma, dres, dgam, dbet, dsum)
    assert rel_l2(dres, want_dx) < 1e-5 and rel_l2(dxb, want_dx) < 5e-3
    assert rel_l2(dgam, gr.grad) < 1e-4 and rel_l2(dbet, br.grad) < 1e-4 and rel_l2(dsum, want_dx.sum(0)) < 1e-4


def test_colsum_cast_sgemm(K):
    a = _bf(3000, 1152, seed=15)
    out = torch.zeros(1152, device="cuda")
    K.colsum_bf16(a, out)
    assert rel_l2(out, a.float().sum(0)) < 1e-5
    src = torch.randn(100003, device="cuda")
    assert torch.equal(K.cast_f32_bf16(src), src.bfloat16())
    A, Bm = torch.randn(45, 70, device="cuda"), torch.randn(70, 161, device="cuda")
    bias = torch.randn(161, device="cuda")
    assert rel_l2(K.sgemm_small(A, Bm, bias=bias), A @ Bm + bias) < 1e-5
    assert rel_l2(K.sgemm_small(A.t().contiguous(), Bm, trans_a=True), A @ Bm) < 1e-5
    assert rel_l2(K.sgemm_small(A, Bm.t().contiguous(), trans_b=True), A @ Bm) < 1e-5
    C = torch.ones(45, 161, device="cuda")
    assert rel_l2(K.sgemm_small(A, Bm, out=C, accumulate=True), A @ Bm + 1) < 1e-5


def test_errors_are_loud(K):
    from diverse_channel_vit_b200._lib import DcvError

    with pytest.raises(DcvError):
        K.gemm_nt(_bf(10, 64), _bf(100, 64))  # N not a multiple of 64
    with pytest.raises(DcvError):
        K.gemm_nt(torch.zeros(8, 64).bfloat16(), torch.zeros(64, 64).bfloat16())  # CPU tensors
