"""Differentiable ops on packed octic rows (torch.autograd.Function wrappers over ops.py).

Dtype policy = what the reference does under torch.autocast(bfloat16): the residual stream and
LayerNorm statistics are fp32, every linear / attention / GELU consumes and produces bf16, accumulation is fp32.
All activations here are 2-D [T, D] (T = B*N tokens) packed rows; parameters are the reference's own fp32 tensors.
"""
from __future__ import annotations

import weakref
from typing import Optional, Tuple

import torch

from . import ops
from ._lib import EPI_BF16, EPI_F32, EPI_GELU_BF16, EPI_GELU_BWD, EPI_RESID, OcticError

# ----------------------------------------------------------------------------------------------------------------
# packing cache: fp32 parameters -> bf16 GEMM operands, re-packed only when a parameter's version changes
# (optimizer steps bump ._version; load_state_dict copies in place and bumps it too)
# ----------------------------------------------------------------------------------------------------------------
_pack_cache: dict = {}
_param_epoch = 0     # bumped by writers that bypass autograd's version counters (optim.FusedOptimizer)


def bump_param_epoch() -> None:
    """Invalidate every cached pack: parameters were rewritten through raw pointers (the fused optimizer kernels)."""
    global _param_epoch
    _param_epoch += 1


def _cached(tensors, build):
    """Cache only for nn.Parameters (stable identity); derived weights (e.g. the expanded patch-embed filters) are
    packed on every call.  An entry is valid while the same objects still hold the same storage at the same version."""
    if not all(isinstance(t, torch.nn.Parameter) for t in tensors):
        with torch.no_grad():
            return build()
    ident = tuple(id(t) for t in tensors)
    key = (_param_epoch,) + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)
    hit = _pack_cache.get(ident)
    if hit is not None and hit[0] == key and all(r() is t for r, t in zip(hit[2], tensors)):
        return hit[1]
    with torch.no_grad():
        pk = build()
    if hit is None:
        # drop the entry (bf16 packs + transposes, ~1.5x the fp32 bytes) as soon as any of its parameters dies, instead of
        # keeping it on the GPU until the id is reused
        for t in tensors:
            weakref.finalize(t, _pack_cache.pop, ident, None)
    _pack_cache[ident] = (key, pk, tuple(weakref.ref(t) for t in tensors))
    return pk


def packed_d8(weights: Tuple[torch.Tensor, ...]) -> ops.PackedD8:
    return _cached(weights, lambda: ops.pack_linear_d8(*[w.detach() for w in weights]))


def packed_dense(weight: torch.Tensor) -> ops.PackedDense:
    return _cached((weight,), lambda: ops.pack_linear(weight.detach().contiguous()))


def clear_pack_cache() -> None:
    _pack_cache.clear()


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------------------------
# gamma-folded backward of  x + gamma * Linear(a)  (layer scale without DropPath; octic_vits/d8_layers.py:698-707,
# deit/vit.py:131-134).  The reference's autograd needs the branch output (saved in forward) for dgamma and makes a
# pass over [T, D] to form dy = bf16(gamma * dres).  Here nothing is saved and no pass is made:
#   * the layer-norm backward that PRODUCES dres also emits g = bf16(dres) and cs = colsum(dres) (two more outputs of
#     a kernel that already holds the row in registers) and parks them in `_aux_slot`;
#   * dgrad:  dx = g (diag(gamma) W)          -- the gamma-scaled transposed packs
#   * wgrad:  dW_raw = g^T a, then one small kernel per layer: dgamma = <W, dW_raw>_rows + b * cs, dW = gamma * dW_raw,
#     db = gamma * cs   (exact: sum_t dres * branch = sum_t dres * (a W^T + b)).
# With DropPath (row_scale) or FOLD_GAMMA = False the branch-saving path (octic_layerscale_bwd) is used.
# ----------------------------------------------------------------------------------------------------------------
FOLD_GAMMA = True
# Weight gradients go straight into an existing fp32 `.grad` (the wgrad kernels accumulate with red.add anyway) and the
# Function returns None for that parameter: no zero fill, no autograd accumulation kernel.  Opt-in PER PARAMETER:
# parallel.FlatGrads tags the parameters whose .grad it owns (`_octic_accumulate`); models it does not manage keep plain
# autograd semantics (tensor hooks see the gradient, torch.autograd.grad has no side effect).  The module flag is a
# process-wide kill switch for tests.
ACCUMULATE_INTO_GRAD = True
ACCUMULATE_ATTR = "_octic_accumulate"
# (data_ptr, shape, bf16 copy, column sums, the tensor itself) of the most recent layer-norm backward output.  One slot on
# purpose: the consumer is the very next autograd node of the same backward; anything else (a second model, a hook that
# replaces the gradient) misses on the address / shape test and takes the recompute path -- the parked tensor is held, so
# its address cannot be recycled while the slot is live.  Measured alternative (round 2, tools/gpu/r2_aw.sh): carrying the
# by-products as an attribute of the gradient tensor loses them in 46 of 64 hand-offs (autograd re-wraps the tensor), i.e.
# +46 streaming passes per step -- not adopted.
_aux_slot = None


def reset_step_state() -> None:
    """Forget every hand-off between autograd nodes of an unfinished step (after a failed CUDA-graph capture)."""
    global _aux_slot
    _aux_slot = None


def _grad_target(p):
    if not ACCUMULATE_INTO_GRAD or not isinstance(p, torch.nn.Parameter) or not getattr(p, ACCUMULATE_ATTR, False):
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape:
        return None
    return g


def _park_aux(dx: torch.Tensor, dxb: torch.Tensor, dxs: torch.Tensor) -> None:
    global _aux_slot
    _aux_slot = (dx.data_ptr(), tuple(dx.shape), dxb, dxs, dx)   # dx itself is held so that its storage cannot be reused


def _take_aux(dres: torch.Tensor):
    """(bf16(dres), colsum(dres)): from the producer when it is the tensor parked last, else one streaming pass."""
    global _aux_slot
    slot, _aux_slot = _aux_slot, None
    if slot is not None and slot[0] == dres.data_ptr() and slot[1] == tuple(dres.shape):
        return slot[2], slot[3]
    g, _, cs = ops.layerscale_bwd(dres, None, None, None, 1, want_colsum=True)
    return g, cs


def packed_d8_scaled_t(weights, gamma, gamma_src):
    """cached on the weight parameters and the parameters gamma was assembled from (gamma itself is a derived tensor)"""
    key = tuple(weights) + tuple(gamma_src if gamma_src is not None else (gamma,))
    return _cached(key, lambda: ops.pack_linear_d8_scaled_t(*[w.detach() for w in weights], gamma.detach().contiguous()))


def packed_dense_scaled_t(weight, gamma):
    return _cached((weight, gamma), lambda: ops.pack_linear_scaled_t(weight.detach().contiguous(), gamma.detach().contiguous()))


# ----------------------------------------------------------------------------------------------------------------
# LinearD8  (reference octic_vits/d8_layers.py:104-127)
# ----------------------------------------------------------------------------------------------------------------
class LinearD8Fn(torch.autograd.Function):
    """y_bf16[T, Dout] = LinearD8(x_bf16[T, Din]).

    head = (H, S) makes the GEMM epilogue write y head-major ([S][H][hd], head vector [A1|A2|B1|B2|E0|E1]) -- the qkv
    operand of the tcgen05 attention (reference pack step d8_layers.py:632-641); backward still receives the packed
    dqkv the attention backward scatters.  dgrad_heads = H makes backward emit dx head-major (proj feeding d_o)."""

    @staticmethod
    def forward(ctx, x, wA1, wA2, wB1, wB2, wE, bias, head=(0, 0), dgrad_heads: int = 0):
        x = _c(x)
        pk = packed_d8((wA1, wA2, wB1, wB2, wE))
        y = torch.empty(x.shape[0], pk.dout, dtype=torch.bfloat16, device=x.device)
        ops.linear_d8(x, pk, bias, EPI_BF16, out=y, head=head)
        ctx.save_for_backward(x)
        ctx.pk = pk
        ctx.has_bias = bias is not None
        ctx.dgrad_heads = dgrad_heads
        ctx.wparams = (wA1, wA2, wB1, wB2, wE)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        pk = ctx.pk
        dy = _c(dy)
        dx = ops.linear_d8_dgrad(dy, pk, ctx.dgrad_heads) if ctx.needs_input_grad[0] else None
        dws = (None,) * 5
        if any(ctx.needs_input_grad[1:6]):
            tg = [_grad_target(w) for w in ctx.wparams]
            if all(t is not None for t in tg):
                ops.linear_d8_wgrad(dy, x, pk.din, pk.dout, targets=tg)
            else:
                dws = ops.linear_d8_wgrad(dy, x, pk.din, pk.dout)
        db = ops.colsum_bf16(dy, pk.dout // 8) if (ctx.has_bias and ctx.needs_input_grad[6]) else None
        return (dx, *dws, db, None, None)


class LinearD8ResidualFn(torch.autograd.Function):
    """resid_out_f32 = resid_f32 + row_scale * gamma * bf16(LinearD8(x_bf16))  -- the proj / fc2 GEMM with the
    LayerScaleD8 / AffineD8(bias=False), DropPathD8 and residual add of the block fused into its epilogue
    (reference d8_layers.py:698-707, 759-776, 205-212, 249-282)."""

    @staticmethod
    def forward(ctx, x, wA1, wA2, wB1, wB2, wE, bias, gamma, resid, row_scale, rows_per_sample, dgrad_heads: int = 0,
                gamma_src=None):
        x = _c(x)
        pk = packed_d8((wA1, wA2, wB1, wB2, wE))
        ctx.fold = FOLD_GAMMA and gamma is not None and row_scale is None and any(ctx.needs_input_grad)
        need_branch = gamma is not None and ctx.needs_input_grad[7] and not ctx.fold
        ctx.gamma_src = gamma_src
        out = torch.empty_like(resid)
        branch = torch.empty(x.shape[0], pk.dout, dtype=torch.bfloat16, device=x.device) if need_branch else None
        ops.linear_d8(x, pk, bias, EPI_RESID, gamma=gamma, resid_in=resid, resid_out=out, row_scale=row_scale,
                      rows_per_sample=rows_per_sample, branch_out=branch)
        ctx.save_for_backward(x, branch, gamma, row_scale, wA1, wA2, wB1, wB2, wE, bias)
        ctx.pk = pk
        ctx.rows_per_sample = rows_per_sample
        ctx.has_bias = bias is not None
        ctx.dgrad_heads = dgrad_heads
        ctx.wparams = (wA1, wA2, wB1, wB2, wE)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, branch, gamma, row_scale, wA1, wA2, wB1, wB2, wE, bias = ctx.saved_tensors
        tg = [_grad_target(w) for w in ctx.wparams]
        fused = all(t is not None for t in tg)
        pk = ctx.pk
        dout = _c(dout)
        if ctx.fold:
            g, cs = _take_aux(dout)
            co = pk.dout // 8
            ws = (wA1, wA2, wB1, wB2, wE)
            dx = None
            if ctx.needs_input_grad[0]:
                dx = ops.linear_d8_dgrad(g, packed_d8_scaled_t(ws, gamma, ctx.gamma_src), ctx.dgrad_heads)
            dws = ops.linear_d8_wgrad(g, x, pk.din, pk.dout)
            dgamma = ops.zeros_f32(gamma.numel(), gamma.device).view_as(gamma)
            db = torch.empty(co, dtype=torch.float32, device=x.device) if ctx.has_bias else None
            gam = gamma.detach()
            segs = [(dws[i], ws[i].detach(), gam[i * co:(i + 1) * co], bias if i == 0 else None,
                     cs[:co] if (i == 0 and ctx.has_bias) else None, dgamma[i * co:(i + 1) * co],
                     db if i == 0 else None, tg[i] if fused else None) for i in range(4)]
            # both E rows share W_E and alpha_E: the whole gradient lands on the first copy of alpha_E in the packed vector
            segs.append((dws[4], wE.detach(), gam[4 * co:6 * co], None, None, dgamma[4 * co:6 * co], None,
                         tg[4] if fused else None))
            ops.layerscale_wgrad_finalize(segs)
            if fused:
                dws = (None,) * 5
            return (dx, *dws, db, dgamma, dout, None, None, None, None)
        dy, dgamma, colsum = ops.layerscale_bwd(dout, branch, gamma, row_scale, ctx.rows_per_sample,
                                                want_colsum=ctx.has_bias)
        dx = ops.linear_d8_dgrad(dy, pk, ctx.dgrad_heads) if ctx.needs_input_grad[0] else None
        dws = (None,) * 5
        if any(ctx.needs_input_grad[1:6]):
            if fused:
                ops.linear_d8_wgrad(dy, x, pk.din, pk.dout, targets=tg)
            else:
                dws = ops.linear_d8_wgrad(dy, x, pk.din, pk.dout)
        db = colsum[: pk.dout // 8] if ctx.has_bias else None
        return (dx, *dws, db, dgamma, dout, None, None, None, None)


# ----------------------------------------------------------------------------------------------------------------
# dense nn.Linear  (deit/vit.py:29-33, timm Mlp)
# ----------------------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = x_bf16 @ W^T + b, bf16 out (or fp32 when out_f32); optional fused exact GELU (saves the pre-activation)."""

    @staticmethod
    def forward(ctx, x, weight, bias, gelu: bool, out_f32: bool):
        x = _c(x)
        pk = packed_dense(weight)
        n, k = pk.n, pk.k
        y = torch.empty(x.shape[0], n, dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
        pre = None
        if gelu:
            pre = torch.empty(x.shape[0], n, dtype=torch.bfloat16, device=x.device) if any(ctx.needs_input_grad) else None
            ops.linear_dense(x, pk.w, n, k, bias, EPI_GELU_BF16, out=y, branch_out=pre)
        else:
            ops.linear_dense(x, pk.w, n, k, bias, EPI_F32 if out_f32 else EPI_BF16, out=y)
        ctx.save_for_backward(x, pre)
        ctx.pk, ctx.gelu, ctx.has_bias = pk, gelu, bias is not None
        ctx.wparam = weight
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pre = ctx.saved_tensors
        pk = ctx.pk
        n, k = pk.n, pk.k
        dy = _c(dy)
        if dy.dtype != torch.bfloat16:
            dy = ops.cast_bf16(dy) if dy.shape[1] % 4 == 0 else dy.to(torch.bfloat16)
        if n % 8:   # TMA needs 16-byte row strides: pad tiny heads (e.g. 10 classes) with zero columns
            dy = torch.nn.functional.pad(dy, (0, 8 - n % 8))
        db = None
        if ctx.gelu:
            colsum = ops.zeros_f32(n, dy.device) if ctx.has_bias else None
            dy = ops.gelu_bwd(dy, pre, colsum)
            db = colsum
        elif ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum_bf16(dy)[:n]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x.shape[0], k, dtype=torch.bfloat16, device=x.device)
            ops.linear_dense(dy, pk.w_t, k, n, None, EPI_BF16, out=dx)
        dw = _dense_wgrad(dy, x, n, k, ctx.wparam) if ctx.needs_input_grad[1] else None
        return dx, dw, db, None, None


class LinearResidualFn(torch.autograd.Function):
    """resid_out = resid + row_scale * gamma * bf16(x @ W^T + b): dense proj / fc2 with layer scale, DropPath and the
    residual add fused (deit/vit.py:131-134).  `remap` scatters GEMM rows into a larger token matrix (patch embed)."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, resid, row_scale, rows_per_sample, remap):
        x = _c(x)
        pk = packed_dense(weight)
        ctx.fold = (FOLD_GAMMA and gamma is not None and row_scale is None and remap == (0, 0, 0)
                    and isinstance(gamma, torch.nn.Parameter) and any(ctx.needs_input_grad))
        need_branch = gamma is not None and ctx.needs_input_grad[3] and not ctx.fold
        out = torch.empty_like(resid)
        if remap != (0, 0, 0):
            out.copy_(resid)        # rows the GEMM does not touch (cls tokens) must carry over
        branch = torch.empty(resid.shape[0], pk.n, dtype=torch.bfloat16, device=x.device) if need_branch else None
        ops.linear_dense(x, pk.w, pk.n, pk.k, bias, EPI_RESID, gamma=gamma, resid_in=resid, resid_out=out,
                         row_scale=row_scale, rows_per_sample=rows_per_sample, branch_out=branch, remap=remap)
        ctx.save_for_backward(x, branch, gamma, row_scale, weight, bias)
        ctx.pk, ctx.rows_per_sample, ctx.has_bias, ctx.remap = pk, rows_per_sample, bias is not None, remap
        ctx.wparam = weight
        return out

    @staticmethod
    def backward(ctx, dout):
        x, branch, gamma, row_scale, weight, bias = ctx.saved_tensors
        pk = ctx.pk
        dout = _c(dout)
        if ctx.fold:
            g, cs = _take_aux(dout)
            dx = None
            if ctx.needs_input_grad[0]:
                dx = torch.empty(x.shape[0], pk.k, dtype=torch.bfloat16, device=x.device)
                ops.linear_dense(g, packed_dense_scaled_t(weight, gamma), pk.k, pk.n, None, EPI_BF16, out=dx)
            dw, dgamma, db = _dense_fold_wgrad(g, x, weight, gamma, bias, cs, pk.n, pk.k, ctx.wparam)
            return dx, dw, db, dgamma, dout, None, None, None
        if ctx.remap != (0, 0, 0):
            grp, extra, off = ctx.remap
            # gather the rows the GEMM wrote: [M/grp, grp + extra, D] -> rows off..off+grp of every group
            d3 = dout.view(-1, grp + extra, dout.shape[1])[:, off:off + grp, :]
            dy = d3.to(torch.bfloat16).reshape(-1, dout.shape[1])
            dgamma = None
            colsum = dy.float().sum(0) if ctx.has_bias else None
        else:
            dy, dgamma, colsum = ops.layerscale_bwd(dout, branch, gamma, row_scale, ctx.rows_per_sample,
                                                    want_colsum=ctx.has_bias)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x.shape[0], pk.k, dtype=torch.bfloat16, device=x.device)
            ops.linear_dense(dy, pk.w_t, pk.k, pk.n, None, EPI_BF16, out=dx)
        dw = _dense_wgrad(dy, x, pk.n, pk.k, ctx.wparam) if ctx.needs_input_grad[1] else None
        return dx, dw, colsum, dgamma, dout, None, None, None


def _dense_fold_wgrad(g, a, weight, gamma, bias, cs, n, k, wparam=None):
    """wgrad + finalize of the gamma-folded backward for one nn.Linear: returns (dW or None if accumulated into
    wparam.grad, dgamma, dbias)."""
    dw = ops.linear_dense_wgrad(g, a, n, k)
    dgamma = ops.zeros_f32(gamma.numel(), gamma.device).view_as(gamma)
    db = torch.empty(n, dtype=torch.float32, device=g.device) if bias is not None else None
    tg = _grad_target(wparam)
    ops.layerscale_wgrad_finalize([(dw, weight.detach(), gamma.detach(), bias, cs if bias is not None else None, dgamma, db,
                                    tg)])
    return (None if tg is not None else dw), dgamma, db


def _dense_wgrad(dy, a, n, k, wparam):
    """plain dense wgrad: accumulated into wparam.grad (returns None) or into a fresh buffer (returned)."""
    tg = _grad_target(wparam)
    if tg is not None:
        ops.linear_dense_wgrad(dy, a, n, k, dw=tg)
        return None
    return ops.linear_dense_wgrad(dy, a, n, k)


class MlpResidualFn(torch.autograd.Function):
    """resid_out = resid + row_scale * gamma * bf16(fc2(gelu(fc1(x)))): the whole dense MLP branch of a block
    (timm Mlp + layer scale + DropPath + residual, deit/vit.py:126-134) as ONE autograd node, so that backward can
    fuse the nn.GELU derivative and the fc1 bias gradient into the epilogue of the fc2 dgrad GEMM
    (OCTIC_EPI_GELU_BWD): the gradient w.r.t. the hidden activation never makes a round trip through HBM."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, resid, row_scale, rows_per_sample):
        x = _c(x)
        pk1, pk2 = packed_dense(w1), packed_dense(w2)
        T = x.shape[0]
        train = any(ctx.needs_input_grad)
        h = torch.empty(T, pk1.n, dtype=torch.bfloat16, device=x.device)
        pre = torch.empty_like(h) if train else None
        ops.linear_dense(x, pk1.w, pk1.n, pk1.k, b1, EPI_GELU_BF16, out=h, branch_out=pre)
        ctx.fold = (FOLD_GAMMA and gamma is not None and row_scale is None and isinstance(gamma, torch.nn.Parameter)
                    and train)
        need_branch = gamma is not None and ctx.needs_input_grad[5] and not ctx.fold
        out = torch.empty_like(resid)
        branch = torch.empty(T, pk2.n, dtype=torch.bfloat16, device=x.device) if need_branch else None
        ops.linear_dense(h, pk2.w, pk2.n, pk2.k, b2, EPI_RESID, gamma=gamma, resid_in=resid, resid_out=out,
                         row_scale=row_scale, rows_per_sample=rows_per_sample, branch_out=branch)
        ctx.save_for_backward(x, pre, h, branch, gamma, row_scale, w2, b2)
        ctx.pk = (pk1, pk2)
        ctx.rows_per_sample = rows_per_sample
        ctx.has_bias = (b1 is not None, b2 is not None)
        ctx.wparams = (w1, w2)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, pre, h, branch, gamma, row_scale, w2, b2 = ctx.saved_tensors
        pk1, pk2 = ctx.pk
        dout = _c(dout)
        db1 = ops.zeros_f32(pk1.n, x.device) if ctx.has_bias[0] else None
        dpre = torch.empty_like(pre)
        if ctx.fold:
            g, cs = _take_aux(dout)
            ops.linear_dense(g, packed_dense_scaled_t(w2, gamma), pk2.k, pk2.n, None, EPI_GELU_BWD, out=dpre, gelu_pre=pre,
                             colsum=db1)
            dw2, dgamma, colsum2 = _dense_fold_wgrad(g, h, w2, gamma, b2, cs, pk2.n, pk2.k, ctx.wparams[1])
        else:
            dy, dgamma, colsum2 = ops.layerscale_bwd(dout, branch, gamma, row_scale, ctx.rows_per_sample,
                                                     want_colsum=ctx.has_bias[1])
            ops.linear_dense(dy, pk2.w_t, pk2.k, pk2.n, None, EPI_GELU_BWD, out=dpre, gelu_pre=pre, colsum=db1)
            dw2 = _dense_wgrad(dy, h, pk2.n, pk2.k, ctx.wparams[1]) if ctx.needs_input_grad[3] else None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x.shape[0], pk1.k, dtype=torch.bfloat16, device=x.device)
            ops.linear_dense(dpre, pk1.w_t, pk1.k, pk1.n, None, EPI_BF16, out=dx)
        dw1 = _dense_wgrad(dpre, x, pk1.n, pk1.k, ctx.wparams[0]) if ctx.needs_input_grad[1] else None
        return dx, dw1, db1, dw2, colsum2, dgamma, dout, None, None


# ----------------------------------------------------------------------------------------------------------------
# packed per-irrep scale vectors  (AffineD8 / LayerScaleD8 parameters, reference d8_layers.py:132-158, 189-212)
# ----------------------------------------------------------------------------------------------------------------
class PackAlphaFn(torch.autograd.Function):
    """(alpha_A1, alpha_A2, alpha_B1, alpha_B2 [C], alpha_E [2C]) -> the packed [8C] vector [A1|A2|B1|B2|E|E] the kernels
    index by packed column (alpha_E serves both E rows).  Backward: the five gradients are slices of the incoming [8C]
    gradient, with the two E halves summed.  When the parameters' .grad tensors are consecutive views of one flat
    buffer (parallel.FlatGrads registers them in this order) they are accumulated with two launches for the whole
    module instead of six slice + accumulate kernels -- an octic block uses four such vectors, 16 blocks per step."""

    @staticmethod
    def forward(ctx, a1, a2, b1, b2, e):
        ctx.params = (a1, a2, b1, b2, e)
        return torch.cat((a1, a2, b1, b2, e, e))

    @staticmethod
    def backward(ctx, g):
        C = ctx.params[0].numel()
        g = _c(g)
        tg = [_grad_target(p) for p in ctx.params]
        if all(t is not None for t in tg):
            base = tg[0]
            offs = (0, C, 2 * C, 3 * C, 4 * C)
            if all(t.data_ptr() == base.data_ptr() + 4 * o for t, o in zip(tg, offs)) and \
                    base.untyped_storage().nbytes() >= 4 * (base.storage_offset() + 6 * C):
                span = torch.as_strided(base, (6 * C,), (1,))
                span.add_(g[:6 * C])
                span[4 * C:].add_(g[6 * C:])
                return None, None, None, None, None
        return g[:C], g[C:2 * C], g[2 * C:3 * C], g[3 * C:4 * C], g[4 * C:6 * C] + g[6 * C:]


# ----------------------------------------------------------------------------------------------------------------
# normalisation
# ----------------------------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    """LayerNormD8 + AffineD8 (d8=True, reference d8_layers.py:161-186, 132-158; alpha is the packed [D] vector with
    alpha_E repeated for both E rows) or nn.LayerNorm (d8=False).  fp32 in, bf16 or fp32 out.

    With passthrough=True the input is returned as a second output: a pre-LN residual block then takes its skip
    connection from that output, so autograd hands both cotangents of x to this backward and the kernel emits
    dx = d_skip + LN^T(dy) in one pass (no separate gradient-accumulation kernel)."""

    @staticmethod
    def forward(ctx, x, alpha, beta, eps: float, d8: bool, out_bf16: bool, passthrough: bool = False):
        if x.stride(-1) != 1:
            x = x.contiguous()
        y, stats = ops.layernorm_fwd(x, alpha, beta, eps, d8, torch.bfloat16 if out_bf16 else torch.float32,
                                     want_stats=any(ctx.needs_input_grad))
        ctx.save_for_backward(x, stats, alpha)
        ctx.d8, ctx.has_beta, ctx.passthrough = d8, beta is not None, passthrough
        if passthrough:
            return y, x.view_as(x)
        return y

    @staticmethod
    def backward(ctx, dy, dskip=None):
        x, stats, alpha = ctx.saved_tensors
        if dy is None:     # only the skip connection was used downstream
            return dskip, None, None, None, None, None, None
        dy = _c(dy)
        if dskip is not None:
            dskip = _c(dskip)
        if FOLD_GAMMA and ctx.passthrough and dskip is not None:
            # x is the output of a residual branch: hand its backward the bf16 copy and the column sums of dx
            dx, dalpha, dbeta, dxb, dxs = ops.layernorm_bwd(dy, x, stats, alpha, ctx.d8, dx_in=dskip, want_aux=True)
            _park_aux(dx, dxb, dxs)
        else:
            dx, dalpha, dbeta = ops.layernorm_bwd(dy, x, stats, alpha, ctx.d8, dx_in=dskip)
        return dx, dalpha, (dbeta if ctx.has_beta else None), None, None, None, None


# ----------------------------------------------------------------------------------------------------------------
# D8 GELU  (reference octic_vits/d8_gelu.py:456-482)
# ----------------------------------------------------------------------------------------------------------------
class GeluD8Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        ctx.save_for_backward(x)
        return ops.gelu_d8_fwd(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ops.gelu_d8_bwd(_c(g), x)


# ----------------------------------------------------------------------------------------------------------------
# attention core  (reference d8_layers.py:632-656 + F.scaled_dot_product_attention; deit/vit.py:36-50)
# ----------------------------------------------------------------------------------------------------------------
class AttentionFn(torch.autograd.Function):
    """layout: ops.ATTN_DENSE, ops.ATTN_OCTIC_PACKED or ops.ATTN_OCTIC_HEADMAJOR (qkv and the incoming d_o head-major,
    o and the outgoing dqkv packed octic rows).
    B, N: ints, or equally long tuples -- the rows are then the concatenation of len(B) segments of B[i] images with
    N[i] tokens each (a DINOv2 crop list: the reference's block-diagonal attention over concatenated crops,
    dinov2/layers/block.py:212-248, is plain attention inside every image); one launch per segment reads / writes
    its row slice in place, every per-token kernel around it runs once over all rows."""

    @staticmethod
    def forward(ctx, qkv, B, N, H: int, hd: int, layout: int):
        qkv = _c(qkv)
        want = any(ctx.needs_input_grad)
        if isinstance(B, int):
            o, lse = ops.attention_fwd(qkv, B, N, H, hd, layout, want_lse=want)
            lses = (lse,)
        else:
            o = torch.empty(qkv.shape[0], H * hd, dtype=torch.bfloat16, device=qkv.device)
            lses, r0 = [], 0
            for b, n in zip(B, N):
                _, lse = ops.attention_fwd(qkv[r0:r0 + b * n], b, n, H, hd, layout, want_lse=want, out=o[r0:r0 + b * n])
                lses.append(lse)
                r0 += b * n
            if r0 != qkv.shape[0]:
                raise OcticError(f"segments cover {r0} rows, qkv has {qkv.shape[0]}")
        ctx.save_for_backward(qkv, o, *[t for t in lses if t is not None])
        ctx.cfg = (B, N, H, hd, layout)
        return o

    @staticmethod
    def backward(ctx, d_o):
        qkv, o, *lses = ctx.saved_tensors
        B, N, H, hd, layout = ctx.cfg
        d_o = _c(d_o)
        if isinstance(B, int):
            return ops.attention_bwd(qkv, o, d_o, lses[0], B, N, H, hd, layout), None, None, None, None, None
        dqkv = torch.empty_like(qkv)
        r0 = 0
        for (b, n), lse in zip(zip(B, N), lses):
            sl = slice(r0, r0 + b * n)
            ops.attention_bwd(qkv[sl], o[sl], d_o[sl], lse, b, n, H, hd, layout, out=dqkv[sl])
            r0 += b * n
        return dqkv, None, None, None, None, None


# ----------------------------------------------------------------------------------------------------------------
# invariantisation / bridge  (reference d8_invariantization.py:49-64, model.py:196-200)
# ----------------------------------------------------------------------------------------------------------------
class PowerSpectrumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        ctx.save_for_backward(x)
        return ops.power_spectrum_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.power_spectrum_bwd(_c(dy), x)


class BridgeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.bridge_permute(_c(x))

    @staticmethod
    def backward(ctx, dy):
        return ops.bridge_permute(_c(dy))


# ----------------------------------------------------------------------------------------------------------------
# symmetric parameter expansion  (reference d8_layers.py:329-373, 475-484; d8_utils.py:388-451)
# ----------------------------------------------------------------------------------------------------------------
def _grad_or_new(p, shape, device):
    """(tensor to write the gradient into, whether it is the parameter's own .grad)"""
    tg = _grad_target(p)
    if tg is not None:
        return tg, True
    return torch.empty(shape, dtype=torch.float32, device=device), False


class LiftWeightFn(torch.autograd.Function):
    """Six half-size filter parameters [Co, Ci, p/2, p/2] -> the GEMM operand [8 Co, Ci p p] of the lifting
    convolution, row blocks A1 | A2 | B1 | B2 | E row 0 = (E_left, E_right) | E row 1 = (rot90 E_left, rot90 E_right).
    maps = {"A1", "A2", "B1", "B2", "E", "Er": ops.SparseMap} over the filter taps."""

    KINDS = ("A1", "A2", "B1", "B2", "E", "E", "Er", "Er")
    SRC = (0, 1, 2, 3, 4, 5, 4, 5)

    @staticmethod
    def forward(ctx, wA1, wA2, wB1, wB2, wEl, wEr, maps):
        ws = (wA1, wA2, wB1, wB2, wEl, wEr)
        Co, Ci, hh, _ = wA1.shape
        taps = 4 * hh * hh
        out = torch.empty(8 * Co, Ci * taps, dtype=torch.float32, device=wA1.device)
        for g, (kind, si) in enumerate(zip(LiftWeightFn.KINDS, LiftWeightFn.SRC)):
            ops.sparse_rowmap(_c(ws[si].detach()).view(Co * Ci, hh * hh), out[g * Co:(g + 1) * Co].view(Co * Ci, taps), maps[kind])
        ctx.maps, ctx.params, ctx.dims = maps, ws, (Co, Ci, hh)
        return out

    @staticmethod
    def backward(ctx, dw):
        Co, Ci, hh = ctx.dims
        taps = 4 * hh * hh
        dw = _c(dw)
        grads, written = [], [False] * 6
        outs = []
        for i, p in enumerate(ctx.params):
            if not ctx.needs_input_grad[i]:
                outs.append(None); grads.append(None)
                continue
            t, own = _grad_or_new(p, p.shape, dw.device)
            outs.append((t, own)); grads.append(None if own else t)
        for g, (kind, si) in enumerate(zip(LiftWeightFn.KINDS, LiftWeightFn.SRC)):
            if outs[si] is None:
                continue
            t, own = outs[si]
            ops.sparse_rowmap(dw[g * Co:(g + 1) * Co].view(Co * Ci, taps), t.view(Co * Ci, hh * hh), ctx.maps[kind], transpose=True,
                              accumulate=own or written[si])
            written[si] = True
        return (*grads, None)


class PosEmbedFn(torch.autograd.Function):
    """Six stored positional-embedding quadrants [h/2, w/2, C] -> packed rows [h w, 8 C]
    (A1 | A2 | B1 | B2 | E row 0 = (x4, x6) | E row 1 = (x5, x7)).  maps over the positions, as in LiftWeightFn."""

    KINDS = ("A1", "A2", "B1", "B2", "E", "E", "Er", "Er")
    SRC = (0, 1, 2, 3, 4, 5, 4, 5)

    @staticmethod
    def forward(ctx, p0, p1, p2, p3, p4, p5, maps):
        ps = (p0, p1, p2, p3, p4, p5)
        h2, w2, C = p0.shape
        n_out = maps["A1"].n_out
        rows = torch.empty(n_out, 8 * C, dtype=torch.float32, device=p0.device)
        for g, (kind, si) in enumerate(zip(PosEmbedFn.KINDS, PosEmbedFn.SRC)):
            ops.sparse_posmap(_c(ps[si].detach()).view(h2 * w2, C), rows[:, g * C:(g + 1) * C], maps[kind])
        ctx.maps, ctx.params = maps, ps
        return rows

    @staticmethod
    def backward(ctx, drows):
        ps = ctx.params
        h2, w2, C = ps[0].shape
        if drows.stride(1) != 1:
            drows = drows.contiguous()
        grads, written, outs = [], [False] * 6, []
        for i, p in enumerate(ps):
            if not ctx.needs_input_grad[i]:
                outs.append(None); grads.append(None)
                continue
            t, own = _grad_or_new(p, p.shape, drows.device)
            outs.append((t, own)); grads.append(None if own else t)
        for g, (kind, si) in enumerate(zip(PosEmbedFn.KINDS, PosEmbedFn.SRC)):
            if outs[si] is None:
                continue
            t, own = outs[si]
            ops.sparse_posmap(drows[:, g * C:(g + 1) * C], t.view(h2 * w2, C), ctx.maps[kind], transpose=True,
                              accumulate=own or written[si])
            written[si] = True
        return (*grads, None)


class Im2ColFn(torch.autograd.Function):
    """Patches of an image as GEMM rows (no gradient w.r.t. the image: it is the network input)."""

    @staticmethod
    def forward(ctx, img, p: int):
        return ops.im2col_patches(_c(img), p)

    @staticmethod
    def backward(ctx, g):
        return None, None


# ----------------------------------------------------------------------------------------------------------------
# 5-tuple <-> packed rows
# ----------------------------------------------------------------------------------------------------------------
def pack_five(xs) -> torch.Tensor:
    """(A1, A2, B1, B2, E) -> [B, N, 8C].  If the five tensors already are the views handed out by `unpack_five`
    of one packed tensor, that tensor is returned without a copy."""
    if len(xs) != 5:
        raise AssertionError("Input should be a 5-tuple")
    a1 = xs[0]
    base = getattr(a1, "_base", None)
    C = a1.shape[-1]
    if (base is not None and base.dim() == 3 and base.shape[-1] == 8 * C and base.is_contiguous()
            and all(getattr(t, "_base", None) is base for t in xs)
            and all(xs[i].storage_offset() == base.storage_offset() + i * C for i in range(4))
            and xs[4].storage_offset() == base.storage_offset() + 4 * C and xs[4].shape[-2:] == (2, 2 * C)):
        return base
    e = xs[4]
    return torch.cat((xs[0], xs[1], xs[2], xs[3], e[..., 0, :], e[..., 1, :]), dim=-1)


def unpack_five(x: torch.Tensor):
    """[B, N, 8C] -> strided views (A1, A2, B1, B2, E[B, N, 2, 2C]) of the same storage."""
    C = x.shape[-1] // 8
    return (x[..., 0:C], x[..., C:2 * C], x[..., 2 * C:3 * C], x[..., 3 * C:4 * C],
            x[..., 4 * C:].unflatten(-1, (2, 2 * C)))


# ----------------------------------------------------------------------------------------------------------------
# torch.compile / autocast contract of the boundary (SURVEY.md section 8b "Threading / streams")
# ----------------------------------------------------------------------------------------------------------------
def opaque_to_compile(fn):
    """Decorator for the forward methods of the reference-facing modules.  The kernels are reached through a C ABI with
    raw pointers; Dynamo cannot (and must not: no compiler-generated kernels on this path) trace into them.  Marking the
    module forwards `torch.compiler.disable` makes `torch.compile(model)` -- what the DeiT recipe does, reference
    deit/main.py:341-342 -- a supported call: Dynamo compiles whatever surrounds these modules and runs them as they
    are, with their hand-written autograd, instead of failing inside ctypes."""
    return torch.compiler.disable(fn, recursive=True)


STRICT_AUTOCAST = False


def set_strict_autocast(on: bool) -> None:
    """The reference decides the GEMM dtype from the autocast context (bf16 under torch.autocast, fp32 without).  This
    implementation has ONE arithmetic: bf16 operands with fp32 accumulation and an fp32 residual stream -- what the
    reference computes under torch.autocast(dtype=torch.bfloat16) -- whether or not an autocast context is active.
    With strict mode on, calling a model outside a CUDA bf16 autocast region raises instead of silently computing in
    bf16 where the reference would have computed in fp32."""
    global STRICT_AUTOCAST
    STRICT_AUTOCAST = bool(on)


def check_autocast() -> None:
    if STRICT_AUTOCAST and not (torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16):
        raise OcticError("strict autocast: this implementation always computes like the reference under "
                         "torch.autocast('cuda', dtype=torch.bfloat16); there is no fp32 (autocast-off) or fp16 GEMM path")


def require_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise OcticError("octic_vits_b200 runs on sm_100 GPUs only: got a CPU tensor (no CPU fallback exists)")
