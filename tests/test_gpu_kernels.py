"""GPU parity tests: every C-ABI kernel against the oracle (oracle/octic_oracle.py) on the same seeded inputs.

The oracle functions are device-agnostic PyTorch; they are evaluated in fp32 (TF32 off) on the bf16-rounded operands
the kernels actually consume, so the comparison isolates the kernel arithmetic.  Tolerances are written per test:
bf16 outputs carry one rounding (2^-9 relative), fp32 outputs of bf16 GEMMs are compared to ~1e-3 of the row scale.
"""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from octic_vits_b200 import ops
    from octic_vits_b200._lib import EPI_BF16, EPI_F32, EPI_GELU_BF16, EPI_GELU_BWD, EPI_RESID
from octic_vits_b200._lib import OcticError
from oracle import octic_oracle as O

DEV = "cuda"
ATTN_DENSE, ATTN_OCTIC_PACKED, ATTN_OCTIC_HEADMAJOR = 0, 1, 2     # include/octic_b200.h OCTIC_ATTN_*


def setup_module(module):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def bf(x):
    return x.to(torch.bfloat16)


def rnd5(B, N, C, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    xs = tuple(torch.randn(B, N, C, generator=g) * scale for _ in range(4)) + (
        torch.randn(B, N, 2, 2 * C, generator=g) * scale,)
    return tuple(x.to(DEV) for x in xs)


def d8_weights(din, dout, bias=True, seed=1, prefix=""):
    g = torch.Generator(device="cpu").manual_seed(seed)
    ci, co = din // 8, dout // 8
    w = {}
    for n in O.IRREPS:
        w[f"{prefix}lin_{n}.weight"] = (torch.randn(co, ci, generator=g) / math.sqrt(ci)).to(DEV)
    w[f"{prefix}lin_E.weight"] = (torch.randn(2 * co, 2 * ci, generator=g) / math.sqrt(2 * ci)).to(DEV)
    if bias:
        w[f"{prefix}lin_A1.bias"] = torch.randn(co, generator=g).to(DEV)
    return w


def pack_d8(w, prefix=""):
    return ops.pack_linear_d8(*(w[f"{prefix}lin_{n}.weight"] for n in ("A1", "A2", "B1", "B2", "E")))


def rounded(w):
    return {k: (bf(v).float() if k.endswith("weight") else v) for k, v in w.items()}


def assert_close(got, want, rtol, atol):
    torch.testing.assert_close(got.float(), want.float(), rtol=rtol, atol=atol)


# ------------------------------------------------------------------------------------------------------------------
# GEMM engine
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,K,N", [(1000, 192, 160), (257, 64, 48), (130, 1280, 1000), (4096, 1280, 3840),
                                   (300, 5120, 1280), (128, 64, 16), (65792 // 8, 1280, 1280)])
def test_gemm_dense_bf16(M, K, N):
    g = torch.Generator().manual_seed(M + K + N)
    x = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    pk = ops.pack_linear(w)
    want = x.float() @ bf(w).float().T + b
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.linear_dense(x, pk.w, N, K, b, EPI_BF16, out=out)
    assert_close(out, want, rtol=1e-2, atol=2e-2)
    out32 = torch.full((M, N), float("nan"), dtype=torch.float32, device=DEV)
    ops.linear_dense(x, pk.w, N, K, b, EPI_F32, out=out32)
    assert_close(out32, want, rtol=1e-4, atol=1e-3)
    # transposed pack: dgrad-style product dy @ W
    dy = bf(torch.randn(M, N, generator=g)).to(DEV)
    dx = torch.empty(M, K, dtype=torch.float32, device=DEV)
    ops.linear_dense(dy, pk.w_t, K, N, None, EPI_F32, out=dx)
    assert_close(dx, dy.float() @ bf(w).float(), rtol=1e-4, atol=1e-3 * math.sqrt(N))


@pytest.mark.parametrize("M,K,N", [(512, 128, 64), (1000, 192, 160), (4099, 1280, 320), (2048, 640, 160)])
def test_gemm_wgrad(M, K, N):
    g = torch.Generator().manual_seed(7 * M + K + N)
    x = bf(torch.randn(M, K, generator=g)).to(DEV)
    dy = bf(torch.randn(M, N, generator=g)).to(DEV)
    dw = ops.linear_dense_wgrad(dy, x, N, K)
    want = dy.float().T @ x.float()
    assert_close(dw, want, rtol=1e-4, atol=2e-3 * math.sqrt(M))
    # accumulation into a pre-loaded buffer and explicit split counts
    dw2 = want.clone()
    ops.linear_dense_wgrad(dy, x, N, K, dw=dw2, splits=3)
    assert_close(dw2, 2 * want, rtol=1e-4, atol=4e-3 * math.sqrt(M))


def test_gemm_gelu_epilogue():
    g = torch.Generator().manual_seed(3)
    M, K, N = 777, 256, 512
    x = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    pk = ops.pack_linear(w, need_t=False)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    pre = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.linear_dense(x, pk.w, N, K, b, EPI_GELU_BF16, out=out, branch_out=pre)
    h = x.float() @ bf(w).float().T + b
    assert_close(pre, h, rtol=1e-2, atol=2e-2)
    assert_close(out, torch.nn.functional.gelu(pre.float()), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("M,K,N", [(777, 256, 512), (300, 192, 1000), (4099, 320, 1280)])
def test_gemm_gelu_bwd_epilogue(M, K, N):
    """fc2 dgrad with the nn.GELU derivative and the fc1 bias gradient fused (OCTIC_EPI_GELU_BWD) vs autograd of
    torch.nn.functional.gelu (deit/vit.py:126-129) on the same bf16 operands."""
    g = torch.Generator().manual_seed(5 + M)
    dy = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = (torch.randn(K, N, generator=g) / math.sqrt(K)).to(DEV)          # fc2 weight [out=K, hidden=N]
    pre = bf(torch.randn(M, N, generator=g) * 1.5).to(DEV)
    pk = ops.pack_linear(w)                                               # w_t = [N, roundup64(K)]
    dpre = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    colsum = torch.zeros(N, dtype=torch.float32, device=DEV)
    ops.linear_dense(dy, pk.w_t, N, K, None, EPI_GELU_BWD, out=dpre, gelu_pre=pre, colsum=colsum)
    p32 = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(p32).backward(dy.float() @ bf(w).float())
    assert_close(dpre, p32.grad, rtol=1e-2, atol=2e-2)
    assert_close(colsum, p32.grad.sum(0), rtol=1e-3, atol=2e-2 * math.sqrt(M))
    with pytest.raises(OcticError):
        ops.linear_dense(dy, pk.w_t, N, K, None, EPI_GELU_BWD, out=dpre)


# ------------------------------------------------------------------------------------------------------------------
# LinearD8
# ------------------------------------------------------------------------------------------------------------------
LIN_CASES = [(3, 259, 384, 1152), (8, 257, 1280, 3840), (2, 197, 1024, 1024), (1, 300, 1280, 5120),
             (1, 300, 5120, 1280), (2, 50, 64, 128)]


@pytest.mark.parametrize("B,N,din,dout", LIN_CASES)
def test_linear_d8_forward(B, N, din, dout):
    xs = rnd5(B, N, din // 8, seed=din + dout)
    w = d8_weights(din, dout)
    x = bf(O.pack_rows(xs)).reshape(B * N, din).contiguous()
    want = O.pack_rows(O.linear_d8(O.unpack_rows(x.float().reshape(B, N, din)), rounded(w), "")).reshape(B * N, dout)
    pk = pack_d8(w)
    out = torch.full((B * N, dout), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.linear_d8(x, pk, w["lin_A1.bias"], EPI_BF16, out=out)
    assert_close(out, want, rtol=1e-2, atol=2e-2)


def test_linear_d8_residual_epilogue():
    B, N, din, dout = 4, 257, 1280, 1280
    xs = rnd5(B, N, din // 8, seed=5)
    w = d8_weights(din, dout)
    x = bf(O.pack_rows(xs)).reshape(B * N, din).contiguous()
    g = torch.Generator().manual_seed(11)
    gamma = torch.randn(dout, generator=g).to(DEV)
    resid = torch.randn(B * N, dout, generator=g).to(DEV)
    keep = torch.tensor([0.0, 2.0, 2.0, 0.0], device=DEV)
    y = O.pack_rows(O.linear_d8(O.unpack_rows(x.float().reshape(B, N, din)), rounded(w), "")).reshape(B * N, dout)
    want = resid + keep.repeat_interleave(N)[:, None] * gamma * bf(y).float()
    pk = pack_d8(w)
    out = torch.empty_like(resid)
    branch = torch.empty(B * N, dout, dtype=torch.bfloat16, device=DEV)
    ops.linear_d8(x, pk, w["lin_A1.bias"], EPI_RESID, gamma=gamma, resid_in=resid, resid_out=out, row_scale=keep,
                  rows_per_sample=N, branch_out=branch)
    assert_close(branch, y, rtol=1e-2, atol=2e-2)
    assert_close(out, want, rtol=1e-2, atol=3e-2)
    # in-place form used at inference
    r2 = resid.clone()
    ops.linear_d8(x, pk, w["lin_A1.bias"], EPI_RESID, gamma=gamma, resid_in=r2, resid_out=r2)
    assert_close(r2, resid + gamma * bf(y).float(), rtol=1e-2, atol=3e-2)


@pytest.mark.parametrize("B,N,din,dout", LIN_CASES)
def test_linear_d8_backward(B, N, din, dout):
    xs = rnd5(B, N, din // 8, seed=1 + din)
    gs = rnd5(B, N, dout // 8, seed=2 + dout)
    w = d8_weights(din, dout)
    x = bf(O.pack_rows(xs)).reshape(B * N, din).contiguous()
    dy = bf(O.pack_rows(gs)).reshape(B * N, dout).contiguous()
    # oracle autograd on the rounded operands
    wr = {k: v.clone().requires_grad_(True) for k, v in rounded(w).items()}
    xin = x.float().reshape(B, N, din).clone().requires_grad_(True)
    y = O.pack_rows(O.linear_d8(O.unpack_rows(xin), wr, ""))
    (y * dy.float().reshape(B, N, dout)).sum().backward()
    pk = pack_d8(w)
    dx = ops.linear_d8_dgrad(dy, pk)
    assert_close(dx, xin.grad.reshape(B * N, din), rtol=1e-2, atol=2e-2 * math.sqrt(dout / 64))
    dws = ops.linear_d8_wgrad(dy, x, din, dout)
    for name, got in zip(("A1", "A2", "B1", "B2", "E"), dws):
        assert_close(got, wr[f"lin_{name}.weight"].grad, rtol=1e-3, atol=2e-3 * math.sqrt(B * N))


# ------------------------------------------------------------------------------------------------------------------
# D8 GELU (reference value test: octic_vits/d8_gelu.py:664-714, rtol = atol = 1e-5 in fp32)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,C", [(16, 196, 32), (8, 196, 64), (3, 257, 640), (2, 5, 6)])
def test_gelu_d8_fp32(B, N, C):
    xs = rnd5(B, N, C, seed=C)
    gs = rnd5(B, N, C, seed=C + 1)
    xin = tuple(x.clone().requires_grad_(True) for x in xs)
    out = O.gelu_d8(xin)
    sum((o * g).sum() for o, g in zip(out, gs)).backward()
    x = O.pack_rows(xs).reshape(B * N, 8 * C).contiguous()
    y = ops.gelu_d8_fwd(x)
    assert_close(y, O.pack_rows(out).reshape(B * N, 8 * C).detach(), rtol=1e-5, atol=1e-5)
    gin = ops.gelu_d8_bwd(O.pack_rows(gs).reshape(B * N, 8 * C).contiguous(), x)
    assert_close(gin, O.pack_rows(tuple(t.grad for t in xin)).reshape(B * N, 8 * C), rtol=1e-5, atol=1e-5)


def test_gelu_d8_bf16_and_bias_colsum():
    B, N, C = 4, 257, 640
    xs = rnd5(B, N, C, seed=9)
    gs = rnd5(B, N, C, seed=10)
    x = bf(O.pack_rows(xs)).reshape(B * N, 8 * C).contiguous()
    g = bf(O.pack_rows(gs)).reshape(B * N, 8 * C).contiguous()
    xin = tuple(t.clone().requires_grad_(True) for t in O.unpack_rows(x.float().reshape(B, N, 8 * C)))
    out = O.gelu_d8(xin)
    sum((o * gg).sum() for o, gg in zip(out, O.unpack_rows(g.float().reshape(B, N, 8 * C)))).backward()
    y = ops.gelu_d8_fwd(x)
    assert_close(y, O.pack_rows(out).reshape(B * N, 8 * C).detach(), rtol=1e-2, atol=1e-2)
    colsum = torch.zeros(C, device=DEV)
    gin = ops.gelu_d8_bwd(g, x, colsum=colsum)
    want = O.pack_rows(tuple(t.grad for t in xin)).reshape(B * N, 8 * C)
    assert_close(gin, want, rtol=1e-2, atol=2e-2)
    assert_close(colsum, gin[:, :C].float().sum(0), rtol=1e-3, atol=5e-2)


# ------------------------------------------------------------------------------------------------------------------
# LayerNormD8 / LayerNorm
# ------------------------------------------------------------------------------------------------------------------
def ln_weights(D, seed=0):
    g = torch.Generator().manual_seed(seed)
    C = D // 8
    w = {f"scaling.alpha_{n}": (1 + 0.3 * torch.randn(C, generator=g)).to(DEV) for n in O.IRREPS}
    w["scaling.alpha_E"] = (1 + 0.3 * torch.randn(2 * C, generator=g)).to(DEV)
    w["scaling.beta"] = (0.3 * torch.randn(C, generator=g)).to(DEV)
    return w


def packed_alpha(w, prefix="scaling."):
    return torch.cat([w[f"{prefix}alpha_{n}"] for n in O.IRREPS] + [w[f"{prefix}alpha_E"]] * 2).contiguous()


@pytest.mark.parametrize("B,N,D", [(4, 257, 1280), (3, 197, 1024), (2, 197, 384), (2, 33, 768), (1, 7, 2048), (1, 5, 64)])
def test_layernorm_d8(B, N, D):
    C = D // 8
    xs = tuple(1.7 * x + 0.4 for x in rnd5(B, N, C, seed=D))
    w = ln_weights(D)
    xin = tuple(x.clone().requires_grad_(True) for x in xs)
    wr = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    out = O.layernorm_d8(xin, wr, "")
    gs = rnd5(B, N, C, seed=D + 1)
    sum((o * g).sum() for o, g in zip(out, gs)).backward()
    x = O.pack_rows(xs).reshape(B * N, D).contiguous()
    alpha = packed_alpha(w)
    y, stats = ops.layernorm_fwd(x, alpha, w["scaling.beta"], 1e-5, d8=True, out_dtype=torch.float32)
    assert_close(y, O.pack_rows(out).reshape(B * N, D).detach(), rtol=1e-5, atol=2e-5)
    ybf, _ = ops.layernorm_fwd(x, alpha, w["scaling.beta"], 1e-5, d8=True, out_dtype=torch.bfloat16)
    assert_close(ybf, y, rtol=1e-2, atol=1e-2)
    dy = O.pack_rows(gs).reshape(B * N, D).contiguous()
    extra = torch.randn(B * N, D, device=DEV)
    dx, dalpha, dbeta = ops.layernorm_bwd(dy, x, stats, alpha, d8=True, dx_in=extra)
    want_dx = O.pack_rows(tuple(t.grad for t in xin)).reshape(B * N, D)
    assert_close(dx - extra, want_dx, rtol=1e-4, atol=1e-4)
    want_da = torch.cat([wr[f"scaling.alpha_{n}"].grad for n in O.IRREPS])
    assert_close(dalpha[:4 * C], want_da, rtol=1e-4, atol=2e-3)
    assert_close(dalpha[4 * C:6 * C] + dalpha[6 * C:], wr["scaling.alpha_E"].grad, rtol=1e-4, atol=2e-3)
    assert_close(dbeta, wr["scaling.beta"].grad, rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("T,D", [(1028, 1280), (591, 1024), (100, 384), (9, 64)])
def test_layernorm_plain(T, D):
    g = torch.Generator().manual_seed(D)
    x = (1.5 * torch.randn(T, D, generator=g) + 0.3).to(DEV)
    w = (1 + 0.3 * torch.randn(D, generator=g)).to(DEV)
    b = (0.3 * torch.randn(D, generator=g)).to(DEV)
    dy = torch.randn(T, D, generator=g).to(DEV)
    xin, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = torch.nn.functional.layer_norm(xin, (D,), wr, br, 1e-6)
    (out * dy).sum().backward()
    y, stats = ops.layernorm_fwd(x, w, b, 1e-6, d8=False, out_dtype=torch.float32)
    assert_close(y, out.detach(), rtol=1e-5, atol=2e-5)
    dx, dw, db = ops.layernorm_bwd(dy, x, stats, w, d8=False)
    assert_close(dx, xin.grad, rtol=1e-4, atol=1e-4)
    assert_close(dw, wr.grad, rtol=1e-4, atol=2e-3)
    assert_close(db, br.grad, rtol=1e-4, atol=2e-3)
    dxb, _, _ = ops.layernorm_bwd(bf(dy), x, stats, w, d8=False)
    assert_close(dxb, xin.grad, rtol=2e-2, atol=2e-2)
    # by-products for the gamma-folded layer-scale backward: bf16 copy and column sums of dx (incl. the skip gradient)
    extra = torch.randn(T, D, device=DEV)
    dx2, dw2, db2, g16, cs = ops.layernorm_bwd(dy, x, stats, w, d8=False, dx_in=extra, want_aux=True)
    assert_close(dx2, xin.grad + extra, rtol=1e-4, atol=1e-4)
    assert_close(dw2, wr.grad, rtol=1e-4, atol=2e-3)
    assert torch.equal(g16, bf(dx2))
    assert_close(cs, dx2.sum(0), rtol=1e-4, atol=2e-3 * math.sqrt(T))


def test_layerscale_bwd():
    B, N, D = 4, 257, 1280
    g = torch.Generator().manual_seed(2)
    dres = torch.randn(B * N, D, generator=g).to(DEV)
    branch = bf(torch.randn(B * N, D, generator=g)).to(DEV)
    gamma = torch.randn(D, generator=g).to(DEV)
    keep = torch.tensor([2.0, 0.0, 2.0, 2.0], device=DEV)
    s = keep.repeat_interleave(N)[:, None]
    dy, dgamma, colsum = ops.layerscale_bwd(dres, branch, gamma, keep, N)
    assert_close(dy, gamma * s * dres, rtol=1e-2, atol=1e-2)
    assert_close(dgamma, (dres * s * branch.float()).sum(0), rtol=1e-4, atol=2e-3)
    assert_close(colsum, dy.float().sum(0), rtol=1e-4, atol=2e-3)
    dy2, dg2, cs2 = ops.layerscale_bwd(dres, None, None, None, 1, want_colsum=False)
    assert dg2 is None and cs2 is None
    assert_close(dy2, dres, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("dense", [True, False])
def test_gamma_folded_layerscale_backward(dense):
    """x + gamma * Linear(a) without DropPath: the gamma-folded backward (scaled dgrad packs + raw wgrad + finalize, no
    saved branch) against torch autograd of the same bf16-operand computation (reference octic_vits/d8_layers.py:698-707,
    deit/vit.py:131-134)."""
    from octic_vits_b200 import functional as OF
    g = torch.Generator().manual_seed(11)
    T, D = 771, 256
    a = bf(torch.randn(T, D, generator=g)).to(DEV)
    resid = torch.randn(T, D, generator=g).to(DEV)
    dres = torch.randn(T, D, generator=g).to(DEV)
    if dense:
        w = torch.nn.Parameter((torch.randn(D, D, generator=g) / math.sqrt(D)).to(DEV))
        b = torch.nn.Parameter(torch.randn(D, generator=g).to(DEV))
        gamma = torch.nn.Parameter((0.5 + torch.rand(D, generator=g)).to(DEV))
        ar = a.clone().requires_grad_(True)
        out = OF.LinearResidualFn.apply(ar, w, b, gamma, resid, None, 1, (0, 0, 0))
        out.backward(dres)
        got = dict(dx=ar.grad, dw=w.grad, db=b.grad, dgamma=gamma.grad)
        a32 = a.float().requires_grad_(True)
        w32, b32, g32 = (t.detach().clone().requires_grad_(True) for t in (w, b, gamma))
        branch = (a32 @ bf(w32).float().T + b32)
        (resid + g32 * branch).backward(dres)
        want = dict(dx=a32.grad, dw=w32.grad, db=b32.grad, dgamma=g32.grad)
    else:
        C = D // 8
        wd = d8_weights(D, D, seed=5)
        names = ["A1", "A2", "B1", "B2", "E"]
        ws = [torch.nn.Parameter(wd[f"lin_{n}.weight"]) for n in names]
        b = torch.nn.Parameter(wd["lin_A1.bias"])
        alphas = [torch.nn.Parameter((0.5 + torch.rand(C if n != "E" else 2 * C, generator=g)).to(DEV)) for n in names]
        gamma = torch.cat(alphas + [alphas[4]])
        ar = a.clone().requires_grad_(True)
        out = OF.LinearD8ResidualFn.apply(ar, *ws, b, gamma, resid, None, 1, 0, tuple(alphas))
        out.backward(dres)
        got = dict(dx=ar.grad, db=b.grad, **{f"dw{n}": w_.grad for n, w_ in zip(names, ws)},
                   **{f"dg{n}": a_.grad for n, a_ in zip(names, alphas)})
        a32 = a.float().requires_grad_(True)
        wr = {f"lin_{n}.weight": bf(w_.detach()).float().requires_grad_(True) for n, w_ in zip(names, ws)}
        wr["lin_A1.bias"] = b.detach().clone().requires_grad_(True)
        al = [a_.detach().clone().requires_grad_(True) for a_ in alphas]
        branch = O.pack_rows(O.linear_d8(O.unpack_rows(a32.reshape(1, T, D)), wr, "")).reshape(T, D)
        (resid + torch.cat(al + [al[4]]) * branch).backward(dres)
        want = dict(dx=a32.grad, db=wr["lin_A1.bias"].grad, **{f"dw{n}": wr[f"lin_{n}.weight"].grad for n in names},
                    **{f"dg{n}": a_.grad for n, a_ in zip(names, al)})
    scale = math.sqrt(T)
    for k in want:
        tol = 3e-2 if k == "dx" else 2e-2 * scale
        assert_close(got[k].float(), want[k], rtol=2e-2, atol=tol)


def test_colsum_and_cast():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3001, 480, generator=g).to(DEV)
    xb = ops.cast_bf16(x)
    assert torch.equal(xb, bf(x))
    assert_close(ops.colsum_bf16(xb), xb.float().sum(0), rtol=1e-4, atol=5e-3)
    assert_close(ops.colsum_bf16(xb, 160), xb[:, :160].float().sum(0), rtol=1e-4, atol=5e-3)


# ------------------------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------------------------
def _attn_reference(qkv_rows, B, N, H, hd, octic):
    D = H * hd
    q3 = qkv_rows.float().reshape(B, N, 3 * D)
    if octic:
        q, k, v = O.attention_heads_d8(O.unpack_rows(q3), H)
    else:
        t = q3.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
        q, k, v = t[0], t[1], t[2]
    o = O.sdpa(q, k, v)
    if octic:
        return O.pack_rows(O.attention_unpack_d8(o)).reshape(B * N, D)
    return o.transpose(1, 2).reshape(B * N, D)


def _o_head_perm(H, hd):
    """perm[c] = head-major column (h*hd + j) that packed octic column c of an attention-output row holds."""
    idx = torch.arange(H * hd, dtype=torch.float32).reshape(1, H, 1, hd)
    return O.pack_rows(O.attention_unpack_d8(idx)).reshape(-1).long()


def _qkv_head_major(qkv_rows, B, N, H, hd):
    """packed LinearD8 qkv rows -> head-major [B*N, 3, H, hd] rows (what the GEMM head remap writes)."""
    q, k, v = O.attention_heads_d8(O.unpack_rows(qkv_rows.float().reshape(B, N, -1)), H)
    return torch.stack((q, k, v), 0).permute(1, 3, 0, 2, 4).reshape(B * N, 3 * H * hd)


ATTN_SHAPES = [(2, 257, 16, 80), (3, 197, 6, 64), (2, 17, 2, 64), (1, 261, 16, 64), (2, 64, 4, 32), (2, 129, 2, 80),
               (2, 128, 2, 96), (1, 400, 2, 80), (1, 150, 2, 128), (3, 65, 2, 64), (5, 1, 2, 64), (2, 272, 2, 80)]


@pytest.mark.parametrize("layout", [ATTN_DENSE, ATTN_OCTIC_PACKED, ATTN_OCTIC_HEADMAJOR])
@pytest.mark.parametrize("B,N,H,hd", ATTN_SHAPES)
def test_attention(B, N, H, hd, layout):
    """Dense and head-major octic layouts run the tcgen05 kernels (attention_tc.cu) inside their envelope -- chunk
    plans with 1..5 chunks, partial last tiles, N = 1 -- and the mma.sync kernels outside it (N = 400 backward,
    hd = 128 backward); the packed octic layout always runs the mma.sync kernels."""
    D = H * hd
    octic = layout != ATTN_DENSE
    if layout == ATTN_OCTIC_HEADMAJOR and not ops.attention_headmajor_ok(N, hd, True):
        pytest.skip("outside the envelope of the tcgen05 kernels: the module falls back to the packed layout")
    g = torch.Generator().manual_seed(N + hd)
    qkv = bf(torch.randn(B * N, 3 * D, generator=g) * 1.5).to(DEV)
    d_o = bf(torch.randn(B * N, D, generator=g)).to(DEV)
    qin = qkv.float().clone().requires_grad_(True)
    want = _attn_reference(qin, B, N, H, hd, octic)
    (want * d_o.float()).sum().backward()
    qkv_in, d_o_in = qkv, d_o
    if layout == ATTN_OCTIC_HEADMAJOR:
        qkv_in = bf(_qkv_head_major(qkv, B, N, H, hd)).contiguous()
        d_o_in = torch.empty_like(d_o)
        d_o_in[:, _o_head_perm(H, hd).to(DEV)] = d_o
    o, lse = ops.attention_fwd(qkv_in, B, N, H, hd, layout)
    assert_close(o, want.detach(), rtol=2e-2, atol=2e-2)
    # log-sum-exp of the scaled scores (natural log), [B, H, N]
    q3 = qkv.float().reshape(B, N, 3 * D)
    if octic:
        qh, kh, _ = O.attention_heads_d8(O.unpack_rows(q3), H)
    else:
        t = q3.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
        qh, kh = t[0], t[1]
    want_lse = torch.logsumexp(qh @ kh.transpose(-1, -2) / hd ** 0.5, dim=-1)
    assert_close(lse, want_lse, rtol=1e-3, atol=1e-3)
    dqkv = ops.attention_bwd(qkv_in, o, d_o_in, lse, B, N, H, hd, layout)
    assert_close(dqkv, qin.grad, rtol=3e-2, atol=3e-2)
    if layout != ATTN_OCTIC_PACKED:
        # the two-pass tcgen05 backward (no scratch: dQ from a recomputed S / dP) must agree with the staged-dQ default
        os.environ["OCTIC_ATTN_STAGED_DQ"] = "0"
        try:
            dqkv2 = ops.attention_bwd(qkv_in, o, d_o_in, lse, B, N, H, hd, layout)
        finally:
            del os.environ["OCTIC_ATTN_STAGED_DQ"]
        assert_close(dqkv2, qin.grad, rtol=3e-2, atol=3e-2)
        assert_close(dqkv2, dqkv, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("growth", [6.0, -6.0])
@pytest.mark.parametrize("B,N,H,hd", [(2, 257, 16, 80), (3, 197, 6, 64), (1, 261, 16, 64), (2, 129, 2, 80)])
def test_attention_extreme_score_ranges(B, N, H, hd, growth):
    """Softmax over score ranges far beyond a comfortable exponent window: keys whose scale grows with the token index
    (growth > 0: every key chunk overshoots the previous ones by far more than 2^8 in the exponent, so the one-pass
    forward's lazy rescale of O and of the row sum runs in EVERY job) or shrinks (growth < 0: the first chunk dominates,
    the rest underflows towards 0).  o, lse and the backward that consumes lse must stay finite and right."""
    D = H * hd
    g = torch.Generator().manual_seed(N * 7 + hd)
    qkv = torch.randn(B, N, 3, H, hd, generator=g) * 1.5
    ramp = torch.linspace(0.0, 1.0, N).view(1, N, 1, 1)
    qkv[:, :, 1] *= torch.exp2(ramp * growth) if growth > 0 else torch.exp2((1.0 - ramp) * -growth)
    qkv = bf(qkv.reshape(B * N, 3 * D)).to(DEV)
    d_o = bf(torch.randn(B * N, D, generator=g)).to(DEV)
    qin = qkv.float().clone().requires_grad_(True)
    want = _attn_reference(qin, B, N, H, hd, False)
    (want * d_o.float()).sum().backward()
    o, lse = ops.attention_fwd(qkv, B, N, H, hd, ATTN_DENSE)
    assert torch.isfinite(o.float()).all() and torch.isfinite(lse).all()
    assert_close(o, want.detach(), rtol=2e-2, atol=2e-2)
    t = qkv.float().reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    want_lse = torch.logsumexp(t[0] @ t[1].transpose(-1, -2) / hd ** 0.5, dim=-1)
    assert_close(lse, want_lse, rtol=1e-3, atol=2e-3)
    dqkv = ops.attention_bwd(qkv, o, d_o, lse, B, N, H, hd, ATTN_DENSE)
    # gradients span six orders of magnitude here (dq = dS K with keys scaled by up to 2^6): norm-wise comparison
    got, ref = dqkv.float(), qin.grad
    assert torch.isfinite(got).all()
    assert float((got - ref).norm() / ref.norm()) < 3e-2
    assert float((got - ref).abs().max() / ref.abs().max()) < 3e-2


@pytest.mark.parametrize("layout", [ATTN_DENSE, ATTN_OCTIC_HEADMAJOR])
def test_attention_bwd_staged_slots_recycle(layout):
    """Staged dQ: 704 CTAs share 2 x SM-count scratch slots (claimed by CAS, released after the last TMA read), launched
    three times back to back on one stream; every launch must reproduce the two-pass kernel's gradients and leave the
    slot flags zero."""
    B, N, H, hd = 44, 257, 16, 80
    D = H * hd
    g = torch.Generator().manual_seed(7)
    qkv = bf(torch.randn(B * N, 3 * D, generator=g)).to(DEV)
    d_o = bf(torch.randn(B * N, D, generator=g)).to(DEV)
    o, lse = ops.attention_fwd(qkv, B, N, H, hd, layout)
    os.environ["OCTIC_ATTN_STAGED_DQ"] = "0"
    try:
        want = ops.attention_bwd(qkv, o, d_o, lse, B, N, H, hd, layout)
    finally:
        del os.environ["OCTIC_ATTN_STAGED_DQ"]
    outs = [ops.attention_bwd(qkv, o, d_o, lse, B, N, H, hd, layout) for _ in range(3)]
    torch.cuda.synchronize()
    for got in outs:
        assert torch.isfinite(got.float()).all()
        assert_close(got, want, rtol=1e-2, atol=1e-2)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    for ws in ops._attn_ws.values():
        assert int(ws[:4096].view(torch.int32).abs().sum()) == 0


def test_attention_headmajor_rejects_unsupported():
    """The head-major layout exists only on the tcgen05 path: outside its envelope the C ABI refuses (no silent gather)."""
    B, N, H, hd = 1, 700, 2, 64
    assert not ops.attention_headmajor_ok(N, hd, False)
    qkv = torch.zeros(B * N, 3 * H * hd, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(OcticError):
        ops.attention_fwd(qkv, B, N, H, hd, ATTN_OCTIC_HEADMAJOR)


@pytest.mark.parametrize("B,N,H,hd", [(2, 257, 16, 80), (3, 197, 6, 64), (1, 33, 2, 32)])
def test_linear_d8_head_remap(B, N, H, hd):
    """qkv LinearD8 with head=(H, 3) writes head-major rows (reference pack step d8_layers.py:632-641 fused into the
    GEMM epilogue); dgrad with head_H writes the proj input gradient head-major.  Same values, permuted columns."""
    D = H * hd
    xs = rnd5(B, N, D // 8, seed=5)
    w = d8_weights(D, 3 * D)
    x = bf(O.pack_rows(xs)).reshape(B * N, D).contiguous()
    pk = pack_d8(w)
    y_packed = torch.empty(B * N, 3 * D, dtype=torch.bfloat16, device=DEV)
    ops.linear_d8(x, pk, w["lin_A1.bias"], EPI_BF16, out=y_packed)
    y_head = torch.empty_like(y_packed)
    ops.linear_d8(x, pk, w["lin_A1.bias"], EPI_BF16, out=y_head, head=(H, 3))
    assert torch.equal(y_head, bf(_qkv_head_major(y_packed, B, N, H, hd)))
    wp = d8_weights(D, D)
    pkp = pack_d8(wp)
    dy = bf(O.pack_rows(rnd5(B, N, D // 8, seed=6))).reshape(B * N, D).contiguous()
    dx_packed = ops.linear_d8_dgrad(dy, pkp)
    dx_head = ops.linear_d8_dgrad(dy, pkp, head_H=H)
    want = torch.empty_like(dx_packed)
    want[:, _o_head_perm(H, hd).to(DEV)] = dx_packed
    assert torch.equal(dx_head, want)


# ------------------------------------------------------------------------------------------------------------------
# invariant / bridge / front end
# ------------------------------------------------------------------------------------------------------------------
def test_power_spectrum_and_bridge():
    B, N, C = 3, 197, 128
    xs = rnd5(B, N, C, seed=3)
    xs[1][0, 0, :4] = 0.0       # exercise sign(0) = 0
    xs[4][0, 0, :, :4] = 0.0    # and the norm subgradient at the origin
    xin = tuple(x.clone().requires_grad_(True) for x in xs)
    y = O.power_spectrum(xin)
    gy = bf(torch.randn(y.shape, device=DEV))
    (y * gy.float()).sum().backward()
    x = O.pack_rows(xs).reshape(B * N, 8 * C).contiguous()
    got = ops.power_spectrum_fwd(x)
    assert_close(got, y.detach().reshape(B * N, 6 * C), rtol=1e-2, atol=1e-2)
    dx = ops.power_spectrum_bwd(gy.reshape(B * N, 6 * C).contiguous(), x)
    assert_close(dx, O.pack_rows(tuple(t.grad for t in xin)).reshape(B * N, 8 * C), rtol=1e-5, atol=1e-5)
    br = ops.bridge_permute(x)
    assert torch.equal(br, O.hybrid_bridge(xs).reshape(B * N, 8 * C))
    assert torch.equal(ops.bridge_permute(br), x)


@pytest.mark.parametrize("p,S", [(14, 224), (16, 224), (8, 32)])
def test_im2col(p, S):
    B = 2
    img = torch.randn(B, 3, S, S, device=DEV)
    got = ops.im2col_patches(img, p)
    K = 3 * p * p
    want = torch.nn.functional.unfold(img, kernel_size=p, stride=p).transpose(1, 2).reshape(-1, K)
    assert torch.equal(got[:, :K], bf(want))
    assert torch.count_nonzero(got[:, K:]) == 0
