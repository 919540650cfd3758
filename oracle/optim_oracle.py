"""CPU oracle for the parameter update of the training step -- TEST INFRASTRUCTURE ONLY (same rules as
oracle/octic_oracle.py: imported by tests/ only, never by the product path).

AdamW  restates torch.optim.AdamW (the DINOv2 recipe's optimizer, reference dinov2/train/train.py:67-68).
       Parity status: PINNED -- tests/test_optim_oracle.py checks it against torch.optim.AdamW itself, which IS the
       reference implementation for that path.
LAMB   restates apex.optimizers.FusedLAMB (apex/optimizers/fused_lamb.py + csrc/multi_tensor_lamb.cu, apex commit
       2386a912164 per the reference's DEIT_ENV.md:5-14), the optimizer the DeiT-III recipe selects
       (experiments/train_deit.py:42 `fusedlamb`, created at deit/main.py:365).  apex is a third-party dependency
       that is neither vendored in the reference nor installed in this image, and the reference holds no test or
       golden vector for it.  Parity status: UNPINNED against apex; written from the published algorithm (You et
       al., "Large Batch Optimization for Deep Learning", 2019) with apex's documented defaults: bias_correction,
       grad_averaging, adam_w_mode (decoupled decay inside the update), global-norm clipping at max_grad_norm = 1.0
       before the moments, trust ratio |w|/|u| only for tensors with weight decay unless use_nvlamb.  The AdamW
       limit (no decay, no clipping) is cross-checked against torch.optim.AdamW.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

Tensor = torch.Tensor


def adamw_step(params: List[Tensor], grads: Sequence[Tensor], exp_avg: List[Tensor], exp_avg_sq: List[Tensor],
               step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decays: Optional[Sequence[float]] = None,
               lr_scales: Optional[Sequence[float]] = None) -> None:
    """torch.optim.AdamW single-tensor rule (torch/optim/adamw.py `_single_tensor_adamw`), in place."""
    b1, b2 = betas
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    for i, (p, g) in enumerate(zip(params, grads)):
        wd = weight_decays[i] if weight_decays is not None else 0.0
        s = lr * (lr_scales[i] if lr_scales is not None else 1.0)
        p.mul_(1 - s * wd)
        exp_avg[i].mul_(b1).add_(g, alpha=1 - b1)
        exp_avg_sq[i].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (exp_avg_sq[i].sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(exp_avg[i], denom, value=-s / bc1)


def lamb_step(params: List[Tensor], grads: Sequence[Tensor], exp_avg: List[Tensor], exp_avg_sq: List[Tensor],
              step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-6, weight_decays: Optional[Sequence[float]] = None,
              lr_scales: Optional[Sequence[float]] = None, max_grad_norm: float = 1.0, grad_averaging: bool = True,
              use_nvlamb: bool = False) -> float:
    """apex FusedLAMB.step (adam_w_mode=True, bias_correction=True), in place; returns the global gradient norm."""
    b1, b2 = betas
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    beta3 = 1 - b1 if grad_averaging else 1.0
    gnorm = math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads))
    clip = gnorm / max_grad_norm if (max_grad_norm > 0 and gnorm > max_grad_norm) else 1.0
    for i, (p, g) in enumerate(zip(params, grads)):
        wd = weight_decays[i] if weight_decays is not None else 0.0
        s = lr * (lr_scales[i] if lr_scales is not None else 1.0)
        sg = g / clip
        exp_avg[i].mul_(b1).add_(sg, alpha=beta3)
        exp_avg_sq[i].mul_(b2).addcmul_(sg, sg, value=1 - b2)
        update = (exp_avg[i] / bc1) / ((exp_avg_sq[i] / bc2).sqrt() + eps) + wd * p
        ratio = s
        if use_nvlamb or wd != 0.0:
            pn, un = float(p.norm()), float(update.norm())
            if pn != 0.0 and un != 0.0:
                ratio = s * pn / un
        p.sub_(update, alpha=ratio)
    return gnorm


def ema_update(ema: List[Tensor], params: Sequence[Tensor], momentum: float) -> None:
    """dinov2/train/ssl_meta_arch.py:370-379 (teacher = m*teacher + (1-m)*student); deit/engine.py:81-82 ModelEma."""
    for e, p in zip(ema, params):
        e.mul_(momentum).add_(p, alpha=1 - momentum)
