"""Pins oracle/optim_oracle.py: AdamW against torch.optim.AdamW (the reference's DINOv2 optimizer), LAMB's AdamW limit
against the same, and the LAMB-specific pieces (clipping, trust ratio) against hand-computed values."""
import math

import pytest
import torch

from oracle import optim_oracle as OO


def make(seed=0, shapes=((5, 7), (7,), (3, 2, 4), (1001,))):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.randn(s, generator=g, dtype=torch.float64) for s in shapes]
    return ps, g


def test_adamw_matches_torch():
    ps, gen = make()
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    wds = [0.05, 0.0, 0.05, 0.0]
    opt = torch.optim.AdamW([{"params": [r], "weight_decay": w} for r, w in zip(ref, wds)], lr=1e-2, betas=(0.9, 0.95),
                            eps=1e-8)
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    for step in range(1, 6):
        grads = [torch.randn(p.shape, generator=gen, dtype=torch.float64) for p in ps]
        for r, g in zip(ref, grads):
            r.grad = g.clone()
        opt.step()
        OO.adamw_step(ps, grads, m, v, step, 1e-2, (0.9, 0.95), 1e-8, wds)
        for a, b in zip(ps, ref):
            torch.testing.assert_close(a, b.detach(), rtol=1e-12, atol=1e-12)


def test_lamb_reduces_to_adamw_without_decay_and_clipping():
    ps, gen = make(1)
    qs = [p.clone() for p in ps]
    m1, v1 = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    m2, v2 = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    for step in range(1, 4):
        grads = [torch.randn(p.shape, generator=gen, dtype=torch.float64) for p in ps]
        OO.lamb_step(ps, grads, m1, v1, step, 3e-3, eps=1e-8, max_grad_norm=0.0)
        OO.adamw_step(qs, grads, m2, v2, step, 3e-3, eps=1e-8)
        for a, b in zip(ps, qs):
            torch.testing.assert_close(a, b, rtol=1e-10, atol=1e-12)


def test_lamb_clipping_and_trust_ratio_by_hand():
    p = [torch.tensor([3.0, 4.0], dtype=torch.float64)]
    g = [torch.tensor([6.0, 8.0], dtype=torch.float64)]            # |g| = 10 -> clipped to (0.6, 0.8)
    m, v = [torch.zeros(2, dtype=torch.float64)], [torch.zeros(2, dtype=torch.float64)]
    gn = OO.lamb_step(p, g, m, v, 1, lr=0.1, betas=(0.9, 0.999), eps=0.0, weight_decays=[0.5], max_grad_norm=1.0)
    assert gn == pytest.approx(10.0)
    # step 1, eps 0: m/bc1 = g', sqrt(v/bc2) = |g'| -> adam part = sign(g') = (1, 1); update = (1, 1) + 0.5*(3, 4)
    upd = torch.tensor([2.5, 3.0], dtype=torch.float64)
    ratio = 0.1 * 5.0 / float(upd.norm())
    torch.testing.assert_close(p[0], torch.tensor([3.0, 4.0], dtype=torch.float64) - ratio * upd)
    # no decay -> no trust ratio (use_nvlamb False)
    p2 = [torch.tensor([3.0, 4.0], dtype=torch.float64)]
    OO.lamb_step(p2, g, [torch.zeros(2, dtype=torch.float64)], [torch.zeros(2, dtype=torch.float64)], 1, lr=0.1, eps=0.0,
                 weight_decays=[0.0])
    torch.testing.assert_close(p2[0], torch.tensor([2.9, 3.9], dtype=torch.float64))
    e = [torch.zeros(2, dtype=torch.float64)]
    OO.ema_update(e, p2, 0.9)
    torch.testing.assert_close(e[0], 0.1 * p2[0])
    assert math.isfinite(float(p[0].sum()))
