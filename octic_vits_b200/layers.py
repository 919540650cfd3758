"""Host-side mirror of the reference's octic layer API (octic_vits/d8_layers.py, d8_invariantization.py, deit/vit.py).

Same class names, constructor signatures, parameter names/shapes (so reference checkpoints load key-for-key) and
error behaviour; the arithmetic runs in liboctic_b200.so.  Every octic module has two entry points:

  forward(xs)          the reference's 5-tuple interface (A1, A2, B1, B2, E); outputs are views of one packed tensor
  forward_packed(x)    packed rows [B, N, D] in, packed rows out -- what the model uses internally (no tuple traffic)

Numerics follow the reference under torch.autocast(bfloat16): fp32 residual stream / LayerNorm, bf16 GEMM I/O.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Tuple, Union

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import functional as OF
from ._lib import OcticError

SQRT2 = math.sqrt(2)
SQRT2_OVER_2 = 0.5 * SQRT2
SQRT2_OVER_4 = 0.5 * SQRT2_OVER_2


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def _rows(x: Tensor) -> Tensor:
    return x.reshape(-1, x.shape[-1])


Segs = Optional[Tuple[Tuple[int, int], ...]]      # ((images, tokens per image), ...) of a concatenated crop list


def _seg_rows(segs) -> int:
    return sum(b * n for b, n in segs)


def _seg_row_scale(per_seg, segs, device) -> Optional[Tensor]:
    """per-segment per-sample residual factors (scale [B_i] or None for each segment) -> one per-ROW vector (the
    residual epilogue is then called with rows_per_sample = 1); None when no segment has one."""
    if all(s is None for s in per_seg):
        return None
    parts = []
    for s, (b, n) in zip(per_seg, segs):
        if s is None:
            s = torch.ones(b, dtype=torch.float32, device=device)
        parts.append(s.repeat_interleave(n))
    return torch.cat(parts)


def _as_bf16(x: Tensor) -> Tensor:
    return x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)


def _as_f32(x: Tensor) -> Tensor:
    return x if x.dtype == torch.float32 else x.float()


# ----------------------------------------------------------------------------------------------------------------
# small modules
# ----------------------------------------------------------------------------------------------------------------
class DropoutD8(nn.Module):
    """reference d8_layers.py:84-96.  p = 0 in every shipped config; p > 0 is applied per tensor like the reference."""

    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.dropout = nn.Dropout(p=p, inplace=inplace)

    def forward(self, xs):
        if self.dropout.p == 0.0 or not self.training:
            return xs
        return tuple(self.dropout(x) for x in xs)


class TritonGeluD8(nn.Module):
    """Drop-in for the reference's TritonGeluD8 (d8_gelu.py:480-482): y = R2I(gelu(I2R(x))), one fused sm_100a kernel
    (not Triton -- the name is kept because the reference passes this class around as `act_layer`)."""

    @OF.opaque_to_compile
    def forward(self, xs):
        x = OF.pack_five(xs)
        OF.require_cuda(x)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        return OF.unpack_five(self.forward_packed(x))

    def forward_packed(self, x: Tensor) -> Tensor:
        return OF.GeluD8Fn.apply(_rows(x)).view(x.shape)


GeluD8 = TritonGeluD8


class LinearD8(nn.Module):
    """reference d8_layers.py:104-130: block-diagonal linear over the irreps; bias on A1 only."""

    def __init__(self, input_channels, output_channels, bias=True):
        super().__init__()
        if input_channels % 8 != 0 or output_channels % 8 != 0:
            raise ValueError()
        self.bias = bias
        self.input_channels = input_channels
        self.output_channels = output_channels
        self.lin_A1 = nn.Linear(input_channels // 8, output_channels // 8, bias=bias)
        self.lin_A2 = nn.Linear(input_channels // 8, output_channels // 8, bias=False)
        self.lin_B1 = nn.Linear(input_channels // 8, output_channels // 8, bias=False)
        self.lin_B2 = nn.Linear(input_channels // 8, output_channels // 8, bias=False)
        self.lin_E = nn.Linear(input_channels // 4, output_channels // 4, bias=False)

    def weights(self):
        return (self.lin_A1.weight, self.lin_A2.weight, self.lin_B1.weight, self.lin_B2.weight, self.lin_E.weight)

    @OF.opaque_to_compile
    def forward(self, x_batched):
        assert len(x_batched) == 5, "Input should be a 5-tuple"
        x = OF.pack_five(x_batched)
        OF.require_cuda(x)
        return OF.unpack_five(self.forward_packed(x))

    def forward_packed(self, x: Tensor, head=(0, 0), dgrad_heads: int = 0) -> Tensor:
        """bf16 (or fp32, cast here) [.., Din] -> bf16 [.., Dout].  head / dgrad_heads: see OF.LinearD8Fn (attention)."""
        y = OF.LinearD8Fn.apply(_as_bf16(_rows(x)), *self.weights(), self.lin_A1.bias, head, dgrad_heads)
        return y.view(*x.shape[:-1], self.output_channels)

    def extra_repr(self) -> str:
        return f"in_features={self.input_channels}, out_features={self.output_channels}, bias={self.bias}"


class AffineD8(nn.Module):
    """reference d8_layers.py:132-158."""

    def __init__(self, dim, bias=True):
        super().__init__()
        if dim % 8 != 0:
            raise ValueError()
        self.alpha_A1 = nn.Parameter(torch.ones(dim // 8))
        self.alpha_A2 = nn.Parameter(torch.ones(dim // 8))
        self.alpha_B1 = nn.Parameter(torch.ones(dim // 8))
        self.alpha_B2 = nn.Parameter(torch.ones(dim // 8))
        self.alpha_E = nn.Parameter(torch.ones(dim // 4))
        self.beta = None
        if bias:
            self.beta = nn.Parameter(torch.zeros(dim // 8))

    def packed_alpha(self) -> Tensor:
        """[D] in packed column order; alpha_E serves both E rows (its gradient sums over them through this cat)."""
        return OF.PackAlphaFn.apply(self.alpha_A1, self.alpha_A2, self.alpha_B1, self.alpha_B2, self.alpha_E)

    def forward(self, xs):
        y0 = self.alpha_A1 * xs[0]
        if self.beta is not None:
            y0 = y0 + self.beta
        return (y0, self.alpha_A2 * xs[1], self.alpha_B1 * xs[2], self.alpha_B2 * xs[3], self.alpha_E * xs[4])


class LayerScaleD8(nn.Module):
    """reference d8_layers.py:189-212."""

    def __init__(self, dim: int, init_values: Union[float, Tensor] = 1e-5) -> None:
        super().__init__()
        if dim % 8 != 0:
            raise ValueError()
        self.alpha_A1 = nn.Parameter(init_values * torch.ones(dim // 8))
        self.alpha_A2 = nn.Parameter(init_values * torch.ones(dim // 8))
        self.alpha_B1 = nn.Parameter(init_values * torch.ones(dim // 8))
        self.alpha_B2 = nn.Parameter(init_values * torch.ones(dim // 8))
        self.alpha_E = nn.Parameter(init_values * torch.ones(dim // 4))

    packed_alpha = AffineD8.packed_alpha

    def forward(self, xs):
        return (self.alpha_A1 * xs[0], self.alpha_A2 * xs[1], self.alpha_B1 * xs[2], self.alpha_B2 * xs[3],
                self.alpha_E * xs[4])


class LayerNormD8(nn.Module):
    """reference d8_layers.py:161-186: six per-irrep means, one shared std, then AffineD8."""

    def __init__(self, channels, eps=1e-05, elementwise_affine=True, bias=True):
        super().__init__()
        self.channels = channels
        self.scaling = AffineD8(channels, bias=bias) if elementwise_affine else nn.Identity()
        self.eps = eps

    def _affine(self, ref: Tensor):
        if isinstance(self.scaling, AffineD8):
            return self.scaling.packed_alpha(), self.scaling.beta
        return torch.ones(self.channels, dtype=torch.float32, device=ref.device), None

    @OF.opaque_to_compile
    def forward(self, xs):
        x = OF.pack_five(xs)
        OF.require_cuda(x)
        return OF.unpack_five(self.forward_packed(_as_f32(x), out_bf16=False))

    def forward_packed(self, x: Tensor, out_bf16: bool = True, passthrough: bool = False):
        alpha, beta = self._affine(x)
        if passthrough:
            y, skip = OF.LayerNormFn.apply(_rows(x), alpha, beta, self.eps, True, out_bf16, True)
            return y.view(x.shape), skip.view(x.shape)
        return OF.LayerNormFn.apply(_rows(x), alpha, beta, self.eps, True, out_bf16).view(x.shape)


def drop_path_d8(xs, drop_prob: float = 0., training: bool = False, scale_by_keep: bool = True):
    """reference d8_layers.py:249-271: one Bernoulli draw per sample shared by all five irrep tensors."""
    if drop_prob == 0. or not training:
        return xs
    s = _drop_scale(xs[0].shape[0], drop_prob, scale_by_keep, xs[0].device)
    return tuple(x * s.to(x.dtype).view(-1, *([1] * (x.dim() - 1))) for x in xs)


def _drop_scale(batch: int, drop_prob: float, scale_by_keep: bool, device) -> Tensor:
    keep = 1.0 - drop_prob
    s = torch.empty(batch, dtype=torch.float32, device=device).bernoulli_(keep)
    if keep > 0.0 and scale_by_keep:
        s.div_(keep)
    return s


class DropPathD8(nn.Module):
    """reference d8_layers.py:273-282."""

    def __init__(self, drop_prob: float = 0., scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def sample(self, batch: int, device) -> Optional[Tensor]:
        """per-sample factor mask/keep_prob (None when inactive); fed to the GEMM epilogue as `row_scale`."""
        if self.drop_prob == 0. or not self.training:
            return None
        return _drop_scale(batch, self.drop_prob, self.scale_by_keep, device)

    def forward(self, xs):
        return drop_path_d8(xs, self.drop_prob, self.training, self.scale_by_keep)


DropPath = DropPathD8  # dense blocks use the same sampler


class MlpD8(nn.Module):
    """reference d8_layers.py:215-247."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=TritonGeluD8, norm_layer=None,
                 bias=True, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        bias = to_2tuple(bias)
        drop_probs = to_2tuple(drop)
        self.fc1 = LinearD8(in_features, hidden_features, bias=bias[0])
        self.act = act_layer()
        self.drop1 = DropoutD8(drop_probs[0])
        self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
        self.fc2 = LinearD8(hidden_features, out_features, bias=bias[1])
        self.drop2 = DropoutD8(drop_probs[1])

    def fusable(self) -> bool:
        return (type(self.act) is TritonGeluD8 and isinstance(self.norm, nn.Identity)
                and self.drop1.dropout.p == 0.0 and self.drop2.dropout.p == 0.0)

    @OF.opaque_to_compile
    def forward(self, xs):
        xs = self.fc1(xs)
        xs = self.act(xs)
        xs = self.drop1(xs)
        xs = self.norm(xs)
        xs = self.fc2(xs)
        xs = self.drop2(xs)
        return xs


class AttentionD8(nn.Module):
    """reference d8_layers.py:590-660.  softmax scale is the SDPA default hd^-1/2 (`scale`/`qk_scale` are dead code in
    the reference and kept here only as attributes)."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = True, proj_bias: bool = True,
                 attn_drop: float = 0.0, proj_drop: float = 0.0, rope=None, qk_scale=None):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        assert (dim // num_heads) % 8 == 0, "dim should be divisible by 8"
        if rope is not None:
            raise NotImplementedError("RoPE not implemented")
        if attn_drop != 0.0:
            raise NotImplementedError("attention dropout is 0 in every reference config; the fused kernel has none")
        self.dim = dim
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = LinearD8(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = LinearD8(dim, dim, bias=proj_bias)
        self.proj_drop = DropoutD8(proj_drop)
        self.rope = rope

    def head_major(self, N: int) -> bool:
        """True when the tcgen05 attention kernels cover (N, head_dim): qkv then leaves its GEMM head-major and the
        proj dgrad returns d_o head-major, so q, k, v, dO reach the kernels by TMA (no gather, no copies)."""
        H = self.num_heads
        return (self.dim // (8 * H)) % 2 == 0 and OF.ops.attention_headmajor_ok(N, self.dim // H)

    def core_packed(self, x: Tensor, segs: Segs = None) -> Tensor:
        """packed bf16/fp32 [B, N, D] -> attention output before `proj`, packed bf16 [B, N, D].  When head_major(N),
        the consumer (`proj`) must be called with dgrad_heads=num_heads.
        segs: x is [1, T, D], the concatenation of segments (B_i images x N_i tokens); every N_i must be head_major."""
        B, N, D = x.shape
        H = self.num_heads
        if segs is not None:
            assert _seg_rows(segs) == B * N and all(self.head_major(n) for _, n in segs)
            qkv = self.qkv.forward_packed(x, head=(H, 3))
            o = OF.AttentionFn.apply(_rows(qkv), tuple(b for b, _ in segs), tuple(n for _, n in segs), H, D // H,
                                     OF.ops.ATTN_OCTIC_HEADMAJOR)
            return o.view(B, N, D)
        if self.head_major(N):
            qkv = self.qkv.forward_packed(x, head=(H, 3))
            o = OF.AttentionFn.apply(_rows(qkv), B, N, H, D // H, OF.ops.ATTN_OCTIC_HEADMAJOR)
        else:
            qkv = self.qkv.forward_packed(x)
            o = OF.AttentionFn.apply(_rows(qkv), B, N, H, D // H, OF.ops.ATTN_OCTIC_PACKED)
        return o.view(B, N, D)

    @OF.opaque_to_compile
    def forward(self, xs):
        x = OF.pack_five(xs)
        OF.require_cuda(x)
        hm = self.head_major(x.shape[1])
        y = self.proj.forward_packed(self.core_packed(x), dgrad_heads=self.num_heads if hm else 0)
        return self.proj_drop(OF.unpack_five(y))


# ----------------------------------------------------------------------------------------------------------------
# octic blocks
# ----------------------------------------------------------------------------------------------------------------
def _branch_residual(lin: LinearD8, a: Tensor, scale_mod, x: Tensor, row_scale: Optional[Tensor],
                     dgrad_heads: int = 0, rows_per_sample: int = 0) -> Tensor:
    """x + row_scale * gamma * lin(a) with everything after the GEMM fused into its epilogue.  rows_per_sample: rows that
    share one row_scale entry (default: the N of x; 1 for the per-row vector of a concatenated crop list)."""
    B, N, D = x.shape
    gamma = scale_mod.packed_alpha() if scale_mod is not None else None
    gamma_src = tuple(scale_mod.parameters()) if scale_mod is not None else None
    out = OF.LinearD8ResidualFn.apply(_as_bf16(_rows(a)), *lin.weights(), lin.lin_A1.bias, gamma, _rows(x), row_scale,
                                      rows_per_sample or N, dgrad_heads, gamma_src)
    return out.view(B, N, D)


class _OcticBlockBase(nn.Module):
    """Shared forward of the two reference block flavours; subclasses name their layer-scale / drop-path children."""

    def _ls(self, i: int):
        raise NotImplementedError

    def _dp(self, i: int):
        raise NotImplementedError

    def _fusable(self) -> bool:
        return (type(self.norm1) is LayerNormD8 and type(self.norm2) is LayerNormD8 and type(self.attn) is AttentionD8
                and type(self.mlp) is MlpD8 and self.mlp.fusable() and self.attn.proj_drop.dropout.p == 0.0
                and all(isinstance(self._ls(i), (AffineD8, LayerScaleD8, nn.Identity)) for i in (1, 2))
                and all(not isinstance(self._ls(i), AffineD8) or self._ls(i).beta is None for i in (1, 2))
                and all(isinstance(self._dp(i), (DropPathD8, nn.Identity)) for i in (1, 2)))

    def supports_segments(self, segs) -> bool:
        """concatenated crop lists need the fused path and the head-major tcgen05 attention for every crop length"""
        return self._fusable() and all(self.attn.head_major(n) for _, n in segs)

    def forward_packed(self, x: Tensor, segs: Segs = None) -> Tensor:
        """fp32 packed [B, N, D] -> fp32 packed [B, N, D]: 8 kernels forward (LN, qkv, attention, proj+ls+dp+res,
        LN, fc1, D8-GELU, fc2+ls+dp+res).
        segs: x is [1, T, D], the concatenated token rows of a crop list (reference NestedTensorBlockD8 list input,
        d8_layers.py:780-794: every crop batch is its own forward, so DropPath draws one factor per image of every
        segment); the per-token kernels run once over all rows, attention once per segment."""
        if not self._fusable():
            assert segs is None
            return OF.pack_five(self._forward_tuple(OF.unpack_five(x)))
        B = x.shape[0]
        ls1 = None if isinstance(self._ls(1), nn.Identity) else self._ls(1)
        ls2 = None if isinstance(self._ls(2), nn.Identity) else self._ls(2)
        rps = 0
        if segs is None:
            s1 = self._dp(1).sample(B, x.device) if isinstance(self._dp(1), DropPathD8) else None
            s2 = self._dp(2).sample(B, x.device) if isinstance(self._dp(2), DropPathD8) else None
            hm = self.attn.head_major(x.shape[1])
        else:
            dp1, dp2 = self._dp(1), self._dp(2)
            # same draw order as crop-by-crop forwards: (branch 1, branch 2) of segment 0, then of segment 1, ..
            draws = [(dp1.sample(b, x.device) if isinstance(dp1, DropPathD8) else None,
                      dp2.sample(b, x.device) if isinstance(dp2, DropPathD8) else None) for b, _ in segs]
            s1 = _seg_row_scale([d[0] for d in draws], segs, x.device)
            s2 = _seg_row_scale([d[1] for d in draws], segs, x.device)
            hm, rps = True, 1
        xn, x = self.norm1.forward_packed(x, passthrough=True)
        a = self.attn.core_packed(xn, segs)
        x = _branch_residual(self.attn.proj, a, ls1, x, s1, dgrad_heads=self.attn.num_heads if hm else 0,
                             rows_per_sample=rps)
        xn, x = self.norm2.forward_packed(x, passthrough=True)
        h = self.mlp.fc1.forward_packed(xn)
        h = self.mlp.act.forward_packed(h)
        return _branch_residual(self.mlp.fc2, h, ls2, x, s2, rows_per_sample=rps)

    @OF.opaque_to_compile
    def forward(self, xs):
        x = OF.pack_five(xs)
        OF.require_cuda(x)
        return OF.unpack_five(self.forward_packed(_as_f32(x)))


class Layer_scale_init_BlockD8(_OcticBlockBase):
    """DeiT-III block, reference d8_layers.py:665-707: x += dp(gamma_1 * attn(norm1 x)); x += dp(gamma_2 * mlp(norm2 x)).
    One DropPathD8 module serves both branches (two independent draws), as in the reference."""

    def __init__(self, dim, num_heads, mlp_ratio=4, qkv_bias=False, qk_scale=None, attn_drop=0., drop=0.,
                 drop_path=0., act_layer=TritonGeluD8, norm_layer=LayerNormD8, Attention_block=AttentionD8,
                 Mlp_block=MlpD8, init_values=1e-4):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention_block(dim=dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                    attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPathD8(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(mlp_ratio * dim)
        self.mlp = Mlp_block(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop)
        self.gamma_1 = AffineD8(dim, bias=False)
        for _, param in self.gamma_1.named_parameters():
            param.data = init_values * torch.ones_like(param.data)
        self.gamma_2 = AffineD8(dim, bias=False)
        for _, param in self.gamma_2.named_parameters():
            param.data = init_values * torch.ones_like(param.data)

    def _ls(self, i):
        return self.gamma_1 if i == 1 else self.gamma_2

    def _dp(self, i):
        return self.drop_path

    def _forward_tuple(self, xs):
        outs = self.drop_path(self.gamma_1(self.attn(self.norm1(xs))))
        xs = tuple(x + o for x, o in zip(xs, outs))
        outs = self.drop_path(self.gamma_2(self.mlp(self.norm2(xs))))
        return tuple(x + o for x, o in zip(xs, outs))


class BlockD8(_OcticBlockBase):
    """DINOv2 / timm-style block, reference d8_layers.py:713-776 (ls1/ls2 = LayerScaleD8 or Identity; drop paths are
    applied only when training and drop_path > 0)."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = False,
                 proj_bias: bool = True, ffn_bias: bool = True, drop: float = 0.0, attn_drop: float = 0.0,
                 init_values=None, drop_path: float = 0.0, act_layer: Callable[..., nn.Module] = TritonGeluD8,
                 norm_layer: Callable[..., nn.Module] = LayerNormD8, attn_class: Callable[..., nn.Module] = AttentionD8,
                 ffn_layer: Callable[..., nn.Module] = MlpD8) -> None:
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = attn_class(dim, num_heads=num_heads, qkv_bias=qkv_bias, proj_bias=proj_bias, attn_drop=attn_drop,
                               proj_drop=drop)
        self.ls1 = LayerScaleD8(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path1 = DropPathD8(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = ffn_layer(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop,
                             bias=ffn_bias)
        self.ls2 = LayerScaleD8(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path2 = DropPathD8(drop_path) if drop_path > 0.0 else nn.Identity()
        self.sample_drop_ratio = drop_path

    def _ls(self, i):
        return self.ls1 if i == 1 else self.ls2

    def _dp(self, i):
        return self.drop_path1 if i == 1 else self.drop_path2

    def _forward_tuple(self, xs):
        residual = self.ls1(self.attn(self.norm1(xs)))
        if self.training and self.sample_drop_ratio > 0.0:
            residual = self.drop_path1(residual)
        xs = tuple(x + r for x, r in zip(xs, residual))
        residual = self.ls2(self.mlp(self.norm2(xs)))
        if self.training and self.sample_drop_ratio > 0.0:
            residual = self.drop_path2(residual)
        return tuple(x + r for x, r in zip(xs, residual))


class NestedTensorBlockD8(BlockD8):
    """reference d8_layers.py:780-794: a tuple, or a list of tuples (one per crop resolution)."""

    def forward_nested(self, x_list: List[Tuple[Tensor]]) -> List[Tuple[Tensor]]:
        return [super(NestedTensorBlockD8, self).forward(x) for x in x_list]

    @OF.opaque_to_compile
    def forward(self, x_or_x_list):
        if isinstance(x_or_x_list, tuple):
            return super().forward(x_or_x_list)
        elif isinstance(x_or_x_list, list):
            return self.forward_nested(x_or_x_list)
        else:
            print(f"Unsupported type: {type(x_or_x_list)}")
            raise AssertionError


# ----------------------------------------------------------------------------------------------------------------
# invariantisation
# ----------------------------------------------------------------------------------------------------------------
class PowerSpectrumInvariant(nn.Module):
    """reference d8_invariantization.py:49-64: cat(A1, |A2|, |B1|, |B2|, ||E||_rows) -> [B, N, 6C] (bf16: it feeds
    invariant_proj, which casts to bf16 under autocast anyway)."""

    def __init__(self, C: int):
        super().__init__()
        self.output_dim = 6 * C // 8

    @OF.opaque_to_compile
    def forward(self, xtuple):
        x = OF.pack_five(tuple(xtuple))
        OF.require_cuda(x)
        return self.forward_packed(_as_f32(x))

    def forward_packed(self, x: Tensor) -> Tensor:
        return OF.PowerSpectrumFn.apply(_rows(x)).view(*x.shape[:-1], self.output_dim)


# ----------------------------------------------------------------------------------------------------------------
# front end
# ----------------------------------------------------------------------------------------------------------------
def expand_filter(weight: Tensor, irrep: str) -> Tensor:
    """The reference's filter symmetrisation (d8_layers.py:329-373) as a function of the stored half-size filter
    [..., p/2, p/2] -> [..., p, p].  "Er" = the E filter turned by 90 degrees (second row of the 2-D irrep, :377-381)."""
    if irrep in ("E", "Er"):
        w = 0.5 * weight
        w2 = torch.cat([w, w.flip(-2)], dim=-2)
        full = torch.cat([w2, -w2.flip(-1)], dim=-1)
        return full.rot90(1, (-2, -1)) if irrep == "Er" else full
    s_rot = 1.0 if irrep in ("A1", "A2") else -1.0
    s_flip = 1.0 if irrep in ("A1", "B1") else -1.0
    w = SQRT2_OVER_4 * weight
    left = torch.cat([w, s_rot * w.rot90(1, (-2, -1))], dim=-2)
    right = torch.cat([s_rot * w.rot90(3, (-2, -1)), w.rot90(2, (-2, -1))], dim=-2)
    full = torch.cat([left, right], dim=-1)
    return full + s_flip * full.flip(-1)


_LIFT_MAPS: dict = {}


def lift_maps(hh: int, device):
    """ops.SparseMap per irrep kind for half-size hh x hh filters, built once per (size, device) by pushing the one-hot
    basis of the stored filter through expand_filter -- the kernel applies exactly the reference formula."""
    key = (hh, str(device))
    if key not in _LIFT_MAPS:
        basis = torch.eye(hh * hh).view(hh * hh, 1, hh, hh)
        _LIFT_MAPS[key] = {k: OF.ops.build_sparse_map(expand_filter(basis, k).reshape(hh * hh, 4 * hh * hh).t().contiguous(), device)
                           for k in ("A1", "A2", "B1", "B2", "E", "Er")}
    return _LIFT_MAPS[key]


class LiftIrrepD8Conv2d(nn.Module):
    """Parameter holder + filter symmetrisation of the reference's lifting convolution (d8_layers.py:284-382).
    The convolution itself is run by PatchEmbedD8 as one im2col GEMM over all irreps."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, bias, irrep="A1"):
        super().__init__()
        if irrep not in ["A1", "A2", "B1", "B2", "E"]:
            raise ValueError("Invalid irrep.")
        if bias and not (irrep == "A1"):
            raise ValueError("Bias only ok for A1-irrep.")
        kernel_size = to_2tuple(kernel_size)
        if kernel_size[0] != kernel_size[1]:
            raise NotImplementedError("Non-square kernels not implemented")
        if kernel_size[0] % 2 != 0 or kernel_size[1] % 2 != 0:
            raise NotImplementedError("Odd kernel sizes not yet implemented")
        if (kernel_size[0] == 2 or kernel_size[1] == 2) and irrep in ["A2", "B1"]:
            raise ValueError(f"No {irrep} irrep in filter kernels of size 2.")
        self.kernel_size = kernel_size
        self.stride = stride
        self.irrep = irrep
        self.bias = None
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size[0] // 2, kernel_size[1] // 2))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            if fan_in != 0:
                bound = 1 / math.sqrt(fan_in)
                nn.init.uniform_(self.bias, -bound, bound)

    def expand_weight(self) -> Tensor:
        """half-size filter -> full D8-symmetric filter [Co, Ci, p, p] (tiny tensors; plain torch, differentiable)."""
        return expand_filter(self.weight, self.irrep)


class LiftD8(nn.Module):
    """reference d8_layers.py:384-411 (children named as in the reference: conv_A1 ... conv_E_right)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, bias):
        super().__init__()
        if out_channels % 8 != 0:
            raise ValueError()
        outs = out_channels // 8
        self.conv_A1 = LiftIrrepD8Conv2d(in_channels, outs, kernel_size, stride, bias=bias, irrep="A1")
        self.conv_A2 = LiftIrrepD8Conv2d(in_channels, outs, kernel_size, stride, bias=False, irrep="A2")
        self.conv_B1 = LiftIrrepD8Conv2d(in_channels, outs, kernel_size, stride, bias=False, irrep="B1")
        self.conv_B2 = LiftIrrepD8Conv2d(in_channels, outs, kernel_size, stride, bias=False, irrep="B2")
        self.conv_E_left = LiftIrrepD8Conv2d(in_channels, outs, kernel_size, stride, bias=False, irrep="E")
        self.conv_E_right = LiftIrrepD8Conv2d(in_channels, outs, kernel_size, stride, bias=False, irrep="E")

    def packed_weight_and_bias(self):
        """GEMM operand [D, Cin*p*p] whose row blocks follow the packed row: A1, A2, B1, B2,
        E row 0 = (E_left, E_right), E row 1 = (rot90 E_left, rot90 E_right) (d8_layers.py:377-381, 475-484)."""
        convs = (self.conv_A1, self.conv_A2, self.conv_B1, self.conv_B2, self.conv_E_left, self.conv_E_right)
        if self.conv_A1.weight.is_cuda and all(type(c) is LiftIrrepD8Conv2d for c in convs):
            # one launch per row block instead of ~150 eager flip / rot90 / cat / mul / add launches (and their backward)
            w = OF.LiftWeightFn.apply(*[c.weight for c in convs], lift_maps(self.conv_A1.weight.shape[-1], self.conv_A1.weight.device))
        else:
            el, er = self.conv_E_left.expand_weight(), self.conv_E_right.expand_weight()
            ws = [self.conv_A1.expand_weight(), self.conv_A2.expand_weight(), self.conv_B1.expand_weight(),
                  self.conv_B2.expand_weight(), el, er, el.rot90(1, (-2, -1)), er.rot90(1, (-2, -1))]
            w = torch.cat([t.flatten(1) for t in ws], dim=0)
        bias = None
        if self.conv_A1.bias is not None:
            bias = torch.cat((self.conv_A1.bias, self.conv_A1.bias.new_zeros(w.shape[0] - self.conv_A1.bias.shape[0])))
        return w, bias


class PatchEmbedD8(nn.Module):
    """reference d8_layers.py:413-497."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True,
                 bias=True, strict_img_size=True):
        super().__init__()
        self.patch_size = to_2tuple(patch_size)
        self.img_size, self.grid_size, self.num_patches = self._init_img_size(img_size)
        if embed_dim % 8 != 0:
            raise ValueError()
        if self.patch_size[0] != self.patch_size[1]:
            raise NotImplementedError("Non-square kernels not implemented")
        self.embed_dim = embed_dim
        self.flatten = flatten
        self.strict_img_size = strict_img_size
        self.lift8 = LiftD8(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size, bias=bias)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def _init_img_size(self, img_size):
        if img_size is None:
            return None, None, None
        img_size = to_2tuple(img_size)
        grid_size = tuple([s // p for s, p in zip(img_size, self.patch_size)])
        return img_size, grid_size, grid_size[0] * grid_size[1]

    def _check(self, x):
        B, C, H, W = x.shape
        if self.img_size is not None:
            if self.strict_img_size:
                torch._assert(H == self.img_size[0], f"Input height ({H}) doesn't match model ({self.img_size[0]}).")
                torch._assert(W == self.img_size[1], f"Input width ({W}) doesn't match model ({self.img_size[1]}).")
            else:
                patch_W, patch_H = self.patch_size
                assert H % (patch_H * 2) == 0, f"Input image height {H} is not an even multiple of patch height {patch_H}"
                assert W % (patch_W * 2) == 0, f"Input image width {W} is not an even multiple of patch width: {patch_W}"

    def embed_into(self, img: Tensor, tokens: Tensor, lead: int) -> Tensor:
        """tokens fp32 [B, lead + np, D] pre-filled with pos-embed / cls rows; adds the patch projections to rows
        lead.. of every image (GEMM epilogue = bias + accumulate into the token matrix).  Returns the new matrix."""
        self._check(img)
        B = img.shape[0]
        p = self.patch_size[0]
        npatch = (img.shape[2] // p) * (img.shape[3] // p)
        w, bias = self.lift8.packed_weight_and_bias()
        patches = OF.Im2ColFn.apply(_as_f32(img), p)
        out = OF.LinearResidualFn.apply(patches, w.contiguous(), bias, None, _rows(tokens), None, 1,
                                        (npatch, lead, lead) if lead else (0, 0, 0))
        return out.view(tokens.shape)

    @OF.opaque_to_compile
    def forward(self, x):
        OF.require_cuda(x)
        self._check(x)
        B = x.shape[0]
        p = self.patch_size[0]
        npatch = (x.shape[2] // p) * (x.shape[3] // p)
        tokens = torch.zeros(B, npatch, self.embed_dim, dtype=torch.float32, device=x.device)
        xs = OF.unpack_five(self.embed_into(x, tokens, 0))
        if not self.flatten:
            raise NotImplementedError("flatten=False is not used by any reference model")
        return self.norm(xs)

    def _init_weights(self):
        for w in [self.lift8.conv_A1.weight.data, self.lift8.conv_A2.weight.data, self.lift8.conv_B1.weight.data,
                  self.lift8.conv_B2.weight.data, self.lift8.conv_E_left.weight.data, self.lift8.conv_E_right.weight.data]:
            torch.nn.init.xavier_uniform_(w.view([w.shape[0], -1]))


class IsotypicToPatchD8(nn.Module):
    """reference d8_layers.py:499-588: octic features -> image patches (the equivariant counterpart of the MAE
    decoder head; exercised by experiments/test_equivariance.py:257-274).  The LinearD8 runs on the grouped tcgen05
    GEMM; the per-token quadrant unfolding is index shuffling on [B, L, p, p, c] and stays in torch."""

    def __init__(self, dim, patch_side, out_channels=3, bias=True, reshape_to_image=False):
        super().__init__()
        if patch_side % 2 != 0:
            raise NotImplementedError("Odd patch side not implemented.")
        self.dim = dim
        self.patch_side = patch_side
        self.out_channels = out_channels
        self.reshape_to_image = reshape_to_image
        self.lin8 = LinearD8(dim, 2 * (patch_side ** 2 * out_channels), bias=bias)

    @staticmethod
    def _quadrants(w: Tensor, s_rot: float, s_flip: float) -> Tensor:
        left = torch.cat((w, s_rot * w.rot90(1, (2, 3))), dim=2)
        right = torch.cat((s_rot * w.rot90(3, (2, 3)), w.rot90(2, (2, 3))), dim=2)
        full = torch.cat((left, right), dim=3)
        return full + s_flip * full.flip(3)

    def forward_packed(self, x: Tensor) -> Tensor:
        B, L, _ = x.shape
        y = self.lin8.forward_packed(x).float()
        h, c = self.patch_side // 2, self.out_channels
        C = y.shape[-1] // 8
        # packed columns: A1 | A2 | B1 | B2 | E row 0 = (x4, x6) | E row 1 = (x5, x7); x6 / x7 are not used (:566-583)
        a1, a2, b1, b2, x4, x5 = (0.25 * y[..., i * C:(i + 1) * C].reshape(B, L, h, h, c) for i in (0, 1, 2, 3, 4, 6))
        out = (self._quadrants(a1, 1.0, 1.0) + self._quadrants(a2, 1.0, -1.0) + self._quadrants(b1, -1.0, 1.0)
               + self._quadrants(b2, -1.0, -1.0))
        for comp, turns in ((x4, 0), (x5, 1)):
            e = SQRT2 * comp
            col = torch.cat((e, e.flip(2)), dim=2)
            full = torch.cat((col, -col.flip(3)), dim=3)
            out = out + (full.rot90(turns, (2, 3)) if turns else full)
        p = self.patch_side
        if self.reshape_to_image:
            H = W = int(math.sqrt(L))
            out = out.reshape(B, H, W, p, p, c)
            return out.permute(0, 5, 1, 3, 2, 4).reshape(B, c, H * p, W * p)
        return out.reshape(B, L, p ** 2 * c)

    def forward(self, xs):
        x = OF.pack_five(xs)
        OF.require_cuda(x)
        return self.forward_packed(x)


# ----------------------------------------------------------------------------------------------------------------
# dense half (deit/vit.py:14-134; timm Block)
# ----------------------------------------------------------------------------------------------------------------
class Attention(nn.Module):
    """reference deit/vit.py:14-56 (fused_attn path; q @ k^T scaled by hd^-1/2)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., fused_attn=True,
                 qk_norm=False, **kwargs):
        super().__init__()
        if attn_drop != 0.0 or proj_drop != 0.0:
            raise NotImplementedError("dropout is 0 in every reference config")
        if qk_scale is not None:
            raise NotImplementedError("custom qk_scale")
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.fused_attn = fused_attn

    def core_packed(self, x: Tensor, segs: Segs = None) -> Tensor:
        B, N, D = x.shape
        qkv = OF.LinearFn.apply(_as_bf16(_rows(x)), self.qkv.weight, self.qkv.bias, False, False)
        if segs is not None:
            assert _seg_rows(segs) == B * N
            return OF.AttentionFn.apply(qkv, tuple(b for b, _ in segs), tuple(n for _, n in segs), self.num_heads,
                                        D // self.num_heads, OF.ops.ATTN_DENSE).view(B, N, D)
        return OF.AttentionFn.apply(qkv, B, N, self.num_heads, D // self.num_heads, OF.ops.ATTN_DENSE).view(B, N, D)

    @OF.opaque_to_compile
    def forward(self, x):
        OF.require_cuda(x)
        B, N, D = x.shape
        o = self.core_packed(x)
        return OF.LinearFn.apply(_rows(o), self.proj.weight, self.proj.bias, False, False).view(B, N, D)


class Mlp(nn.Module):
    """timm Mlp as used by deit/vit.py:9,126-129 (fc1 -> GELU -> fc2, dropout 0)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None,
                 bias=True, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if drop != 0.0:
            raise NotImplementedError("dropout is 0 in every reference config")
        if act_layer is not nn.GELU:
            raise NotImplementedError("the fused fc1 epilogue implements exact (erf) GELU only")
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def hidden_packed(self, x: Tensor) -> Tensor:
        return OF.LinearFn.apply(_as_bf16(_rows(x)), self.fc1.weight, self.fc1.bias, True, False)

    @OF.opaque_to_compile
    def forward(self, x):
        OF.require_cuda(x)
        h = self.hidden_packed(x)
        y = OF.LinearFn.apply(h, self.fc2.weight, self.fc2.bias, False, False)
        return y.view(*x.shape[:-1], y.shape[-1])


class _DenseBlockBase(nn.Module):
    def _gamma(self, i):
        raise NotImplementedError

    def _dp(self, i):
        raise NotImplementedError

    def _drop_scales(self, batch: int, device, nested: bool = False):
        """per-sample residual factors (mask / keep) of the two branches, None when stochastic depth is inactive"""
        s1 = self._dp(1).sample(batch, device) if isinstance(self._dp(1), DropPathD8) else None
        s2 = self._dp(2).sample(batch, device) if isinstance(self._dp(2), DropPathD8) else None
        return s1, s2

    @OF.opaque_to_compile
    def forward(self, x, nested: bool = False, segs: Segs = None):
        """fp32 [B, N, D] -> fp32 [B, N, D]; 7 kernels: LN, qkv, attention, proj+ls+dp+res, LN, fc1+GELU, fc2+ls+dp+res.
        `nested`: the input is one element of a crop list (only the DINOv2 block's stochastic-depth rule cares).
        `segs`: x is [1, T, D], the concatenated rows of a whole crop list (see _OcticBlockBase.forward_packed)."""
        OF.require_cuda(x)
        x = _as_f32(x)
        B, N, D = x.shape
        if segs is None:
            s1, s2 = self._drop_scales(B, x.device, nested)
        else:
            draws = [self._drop_scales(b, x.device, True) for b, _ in segs]
            s1 = _seg_row_scale([d[0] for d in draws], segs, x.device)
            s2 = _seg_row_scale([d[1] for d in draws], segs, x.device)
            N = 1 if (s1 is not None or s2 is not None) else N      # rows per row_scale entry
        xn, xs = OF.LayerNormFn.apply(_rows(x), self.norm1.weight, self.norm1.bias, self.norm1.eps, False, True, True)
        a = self.attn.core_packed(xn.view(x.shape), segs) if segs is not None else self.attn.core_packed(xn.view(B, N, D))
        x2 = OF.LinearResidualFn.apply(_rows(a), self.attn.proj.weight, self.attn.proj.bias, self._gamma(1), xs,
                                       s1, N, (0, 0, 0))
        xn, x2 = OF.LayerNormFn.apply(x2, self.norm2.weight, self.norm2.bias, self.norm2.eps, False, True, True)
        if type(self.mlp) is Mlp and isinstance(self.mlp.norm, nn.Identity):
            out = OF.MlpResidualFn.apply(_as_bf16(xn), self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight,
                                         self.mlp.fc2.bias, self._gamma(2), x2, s2, N)
        else:
            h = self.mlp.hidden_packed(xn)
            out = OF.LinearResidualFn.apply(h, self.mlp.fc2.weight, self.mlp.fc2.bias, self._gamma(2), x2, s2, N,
                                            (0, 0, 0))
        return out.view(x.shape)


class Layer_scale_init_Block(_DenseBlockBase):
    """reference deit/vit.py:90-134 (dense half of the DeiT-III hybrids): gamma_1 / gamma_2 are plain [D] parameters."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, Attention_block=Attention, Mlp_block=Mlp,
                 init_values=1e-4, use_fused_attn=True):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention_block(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                    proj_drop=drop, fused_attn=use_fused_attn)
        self.drop_path = DropPathD8(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = Mlp_block(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop)
        self.gamma_1 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
        self.gamma_2 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)

    def _gamma(self, i):
        return self.gamma_1 if i == 1 else self.gamma_2

    def _dp(self, i):
        return self.drop_path


class LayerScale(nn.Module):
    """timm LayerScale (parameter name `gamma`)."""

    def __init__(self, dim, init_values=1e-5, inplace=False):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class Block(_DenseBlockBase):
    """timm.models.vision_transformer.Block (timm 1.0.x child names), the default `standard_block_layers` of the
    reference model (model.py:21,63)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_norm=False, proj_drop=0., attn_drop=0.,
                 init_values=None, drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, mlp_layer=Mlp, **kwargs):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=proj_drop)
        self.ls1 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path1 = DropPathD8(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = mlp_layer(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=proj_drop)
        self.ls2 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path2 = DropPathD8(drop_path) if drop_path > 0. else nn.Identity()

    def _gamma(self, i):
        ls = self.ls1 if i == 1 else self.ls2
        return ls.gamma if isinstance(ls, LayerScale) else None

    def _dp(self, i):
        return self.drop_path1 if i == 1 else self.drop_path2
