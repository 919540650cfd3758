"""DINOv2 backbone of the reference (octic_vits/dinov2_models.py:40-329) on the packed-row sm_100a kernels.

Same class / factory names, constructor arguments, state-dict keys and output dictionary as the reference:

  OcticDinoVisionTransformer(img_size, patch_size, embed_dim, depth, num_heads, mlp_ratio, num_register_tokens,
                             drop_path_rate, octic_block_layers, standard_block_layers, invariant, **kwargs)
  forward(x | [x...], masks=None | [masks...], is_training=False)
  forward_features / forward_features_list / prepare_tokens_with_masks / get_intermediate_layers / no_weight_decay

and the dense half the reference takes from dinov2/layers (block.py:43-261 `Block` / `NestedTensorBlock`,
attention.py:36-89 `Attention` / `MemEffAttention`, mlp.py:16-40 `Mlp`, layer_scale.py `LayerScale`).

Differences from the reference, all of them reference defects that this mirror does not
inherit: list inputs (multi-crop) run without xformers -- the reference's block-diagonal attention over concatenated
crops is, per crop, plain attention within each image, which is what the kernels compute crop by crop;
`get_intermediate_layers(reshape=True)` works (the reference divides by the patch-size tuple); an explicit
`octic_equi_break_layer=k` is honoured (the reference subclass swallows it and always uses depth/2, which stays the
default).  Like the reference, `init_values=init_scale` (1e-4) reaches every block, overriding the factories'
`partial(..., init_values=1e-5)` (model.py:116-137).
"""
from __future__ import annotations

from functools import partial
from typing import Callable, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from . import functional as OF
from .deit_models import register_model
from .layers import (SQRT2_OVER_2, Attention, DropPathD8, LayerScale, Mlp, NestedTensorBlockD8, _DenseBlockBase, _rows)
from .model import OcticVisionTransformer, trunc_normal_, unfold_pos_embed_packed

# packed-row column block of 8-tuple component k (A1, A2, B1, B2, E11, E21, E12, E22): E row 0 = (x4, x6), row 1 = (x5, x7)
_PACK_ORDER = (0, 1, 2, 3, 4, 6, 5, 7)


def named_apply(fn: Callable, module: nn.Module, name="", depth_first=True, include_root=False) -> nn.Module:
    """Same contract as the reference helper (dinov2_models.py:18-26): call fn(module=, name=) on every descendant
    (and on the root when include_root), children before parents when depth_first."""
    if depth_first:                       # post-order: a module after all of its descendants
        visits = _post_order(module, name, include_root)
    else:                                 # pre-order, which is what named_modules() yields
        visits = [(n if not name else (f"{name}.{n}" if n else name), m) for n, m in module.named_modules()]
        if not include_root:
            visits = visits[1:]
    for n, m in visits:
        fn(module=m, name=n)
    return module


def _post_order(module: nn.Module, name: str, include_self: bool):
    out = []
    for child_name, child in module.named_children():
        out += _post_order(child, f"{name}.{child_name}" if name else child_name, True)
    if include_self:
        out.append((name, module))
    return out


def init_weights_vit_timm(module: nn.Module, name: str = ""):
    """reference dinov2_models.py:28-33"""
    if isinstance(module, nn.Linear):
        trunc_normal_(module.weight, std=0.02)
        if module.bias is not None:
            nn.init.zeros_(module.bias)


# ----------------------------------------------------------------------------------------------------------------
# dense half: dinov2/layers
# ----------------------------------------------------------------------------------------------------------------
class MemEffAttention(Attention):
    """dinov2/layers/attention.py:36-89 (`Attention` and `MemEffAttention` compute the same thing; the xformers call
    is replaced by the tcgen05 attention kernel).  `attn_bias` exists for signature parity: lists of crops are handled
    by the block, so a non-None bias is an error exactly as in the reference without xformers (:73-76)."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = False, proj_bias: bool = True,
                 attn_drop: float = 0.0, proj_drop: float = 0.0) -> None:
        super().__init__(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=proj_drop)
        if not proj_bias:
            self.proj = nn.Linear(dim, dim, bias=False)

    def forward(self, x: Tensor, attn_bias=None) -> Tensor:
        if attn_bias is not None:
            raise AssertionError("xFormers is required for using nested tensors")
        return super().forward(x)


def subset_drop_scale(batch: int, sample_drop_ratio: float, device) -> Tensor:
    """drop_add_residual_stochastic_depth (dinov2/layers/block.py:117-140) as a per-sample residual factor: the
    reference evaluates the branch on a random subset of max(int(b*(1-ratio)), 1) samples (`torch.randperm(b)[:subset]`)
    and index_adds it back with alpha = b/subset; that equals  x + s * branch(x)  with s = b/subset on the subset and 0
    elsewhere.  The uniformly random subset is drawn as "the `keep` smallest of b iid uniforms" with elementwise ops only
    (rank by an all-pairs comparison, b <= a few hundred): no sort, no host synchronisation, so the draw can be
    captured in a CUDA graph and every replay draws a fresh subset (parallel.GraphedStep)."""
    keep = max(int(batch * (1 - sample_drop_ratio)), 1)
    r = torch.rand(batch, device=device)
    idx = torch.arange(batch, device=device)
    # rank of sample i = number of samples with a smaller draw (ties, probability ~0, broken by index)
    before = (r[None, :] < r[:, None]) | ((r[None, :] == r[:, None]) & (idx[None, :] < idx[:, None]))
    rank = before.sum(dim=1)
    return (rank < keep).to(torch.float32) * (batch / keep)


class Block(_DenseBlockBase):
    """dinov2/layers/block.py:43-114: pre-LN block with LayerScale and batch-subset stochastic depth."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = False,
                 proj_bias: bool = True, ffn_bias: bool = True, drop: float = 0.0, attn_drop: float = 0.0,
                 init_values=None, drop_path: float = 0.0, act_layer: Callable[..., nn.Module] = nn.GELU,
                 norm_layer: Callable[..., nn.Module] = nn.LayerNorm, attn_class: Callable[..., nn.Module] = MemEffAttention,
                 ffn_layer: Callable[..., nn.Module] = Mlp) -> None:
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = attn_class(dim, num_heads=num_heads, qkv_bias=qkv_bias, proj_bias=proj_bias, attn_drop=attn_drop,
                               proj_drop=drop)
        self.ls1 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path1 = DropPathD8(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = ffn_layer(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop,
                             bias=ffn_bias)
        self.ls2 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path2 = DropPathD8(drop_path) if drop_path > 0.0 else nn.Identity()
        self.sample_drop_ratio = drop_path

    def _gamma(self, i):
        ls = self.ls1 if i == 1 else self.ls2
        return ls.gamma if isinstance(ls, LayerScale) else None

    def _dp(self, i):
        return self.drop_path1      # the reference uses drop_path1 for both branches (block.py:107-108 "FIXME")

    def _drop_scales(self, batch: int, device, nested: bool = False):
        """block.py:92-112 (tensor input: subset sampling above ratio 0.1, Bernoulli DropPath below) and :219-240
        (list input: subset sampling for any ratio > 0)."""
        r = self.sample_drop_ratio
        if not self.training or r <= 0.0:
            return None, None
        if nested or r > 0.1:
            return subset_drop_scale(batch, r, device), subset_drop_scale(batch, r, device)
        return self.drop_path1.sample(batch, device), self.drop_path1.sample(batch, device)


class NestedTensorBlock(Block):
    """dinov2/layers/block.py:211-261: a tensor, or a list of tensors (one per crop resolution).  Each list element is
    processed as its own batch: attention never crosses images, and every other op is per token."""

    def forward_nested(self, x_list: List[Tensor]) -> List[Tensor]:
        assert isinstance(self.attn, MemEffAttention)
        return [_DenseBlockBase.forward(self, x, nested=True) for x in x_list]

    @OF.opaque_to_compile
    def forward(self, x_or_x_list):
        if isinstance(x_or_x_list, Tensor):
            return super().forward(x_or_x_list)
        elif isinstance(x_or_x_list, list):
            return self.forward_nested(x_or_x_list)
        else:
            print(f"Unsupported type: {type(x_or_x_list)}")
            raise AssertionError


class BlockChunk(nn.ModuleList):
    """reference dinov2_models.py:34-38"""

    def forward(self, x):
        for b in self:
            x = b(x)
        return x


# ----------------------------------------------------------------------------------------------------------------
# the backbone
# ----------------------------------------------------------------------------------------------------------------
class OcticDinoVisionTransformer(OcticVisionTransformer):
    """reference dinov2_models.py:40-260."""

    def __init__(self, img_size: int = 224, patch_size: int = 16, embed_dim: int = 768, depth: int = 12,
                 num_heads: int = 12, mlp_ratio: float = 4.0, num_register_tokens: int = 0,
                 drop_path_rate: float = 0.0, octic_block_layers: Callable = NestedTensorBlockD8,
                 standard_block_layers: Callable = partial(NestedTensorBlock, attn_class=MemEffAttention),
                 invariant: bool = False, octic_equi_break_layer: Optional[int] = None,
                 dynamic_img_size: bool = False, **kwargs):
        """`octic_equi_break_layer` = number of leading octic blocks (README.md:36).  The reference subclass swallows
        it in **kwargs and always breaks at depth/2; None keeps that behaviour, an int is
        honoured here because BASELINE.json configs[4] sweeps it.  `dynamic_img_size=True` (not in the reference)
        enables the multi-resolution path the reference intends but cannot run (Appendix B.1): square inputs whose side
        is an even multiple of the patch size get a bicubically interpolated pos-embed (`interpolate_spatial_tuple`,
        d8_utils.py:453-499) -- this is what DINOv2 local crops need.  Other **kwargs are ignored as in the reference."""
        super().__init__(img_size=img_size, patch_size=patch_size, embed_dim=embed_dim, depth=depth,
                         num_heads=num_heads, mlp_ratio=mlp_ratio, num_register_tokens=num_register_tokens,
                         octic_block_layers=octic_block_layers, standard_block_layers=standard_block_layers,
                         drop_path_rate=drop_path_rate, invariant=invariant, qkv_bias=True, ffn_bias=True,
                         proj_bias=True, octic_equi_break_layer=octic_equi_break_layer)
        self.depth = depth
        self.dynamic_img_size = dynamic_img_size
        if dynamic_img_size:
            self.patch_embed.strict_img_size = False
        C = embed_dim // 8
        g = img_size // patch_size // 2
        self.cls_token = nn.ParameterList([nn.Parameter(torch.zeros(1, 1, C), requires_grad=(i == 0)) for i in range(8)])
        self.pos_embed = nn.ParameterList([nn.Parameter(torch.empty(g, g, C)) for _ in range(6)])
        assert num_register_tokens >= 0
        self.register_tokens = (
            nn.ParameterList([nn.Parameter(torch.zeros(1, num_register_tokens, C), requires_grad=(i == 0))
                              for i in range(8)]) if num_register_tokens else None)
        assert depth % 2 == 0, "depth should be even!"
        self.chunked_blocks = False
        self.mask_token = nn.ParameterList([nn.Parameter(torch.zeros(1, C), requires_grad=(i == 0)) for i in range(8)])
        self.head = nn.Identity()
        self.init_weights()

    def init_weights(self):
        std = 8 * 0.02
        for p in list(self.pos_embed):
            trunc_normal_(p, std=std * SQRT2_OVER_2)
        nn.init.normal_(self.cls_token[0], std=1e-6)
        if self.register_tokens is not None:
            nn.init.normal_(self.register_tokens[0], std=1e-6)
        named_apply(init_weights_vit_timm, self)

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _packed(plist) -> Tensor:
        """8 per-component parameters [.., C] -> packed columns [.., D]"""
        return torch.cat([plist[k] for k in _PACK_ORDER], dim=-1)

    def prepare_tokens_packed(self, x: Tensor, masks: Optional[Tensor] = None) -> Tensor:
        """prepare_tokens_with_masks (reference :113-136) -> fp32 packed [B, 1 + R + np, D]: rows are pre-filled with
        cls | registers | pos-embed, the patch GEMM accumulates into the patch rows, masked patches are replaced by
        mask_token + pos-embed."""
        OF.require_cuda(x)
        B = x.shape[0]
        pos = unfold_pos_embed_packed(self.pos_embed)                                   # [np, D]
        p = self.patch_embed.patch_size[0]
        h0, w0 = x.shape[-2] // p, x.shape[-1] // p
        if pos.shape[0] != h0 * w0 or h0 != w0:
            if not self.dynamic_img_size:
                raise NotImplementedError("positional-embedding interpolation needs dynamic_img_size=True (the "
                                          "reference path is broken, SURVEY Appendix B.1)")
            M = int(round(pos.shape[0] ** 0.5))
            grid = pos.view(M, M, -1).permute(2, 0, 1).unsqueeze(0).float()
            grid = nn.functional.interpolate(grid, size=(h0, w0), mode="bicubic", antialias=False)
            pos = grid[0].permute(1, 2, 0).reshape(h0 * w0, -1)
        lead_rows = [self._packed(self.cls_token)[0]]
        if self.register_tokens is not None:
            lead_rows.append(self._packed(self.register_tokens)[0])
        lead = sum(r.shape[0] for r in lead_rows)
        rows = torch.cat(lead_rows + [pos], dim=0)
        tokens = rows.unsqueeze(0).expand(B, -1, -1).contiguous()
        t = self.patch_embed.embed_into(x, tokens, lead)
        if masks is not None:
            masked_rows = self._packed(self.mask_token) + pos                           # [np, D]
            full = torch.cat((masks.new_zeros(B, lead), masks.to(torch.bool)), dim=1)
            t = torch.where(full.unsqueeze(-1), torch.cat((rows[:lead], masked_rows), dim=0).unsqueeze(0), t)
        return t

    def prepare_tokens_with_masks(self, x, masks=None):
        """reference interface: the 5-tuple (views of the packed matrix)"""
        return OF.unpack_five(self.prepare_tokens_packed(x, masks))

    def _backbone_packed(self, t: Tensor, take: Sequence[int] = ()) -> Tuple[Tensor, List[Tensor]]:
        """octic blocks -> bridge / invariantisation -> dense blocks on one crop batch; returns the pre-norm tokens
        and the outputs of the dense blocks listed in `take`."""
        half = self.octic_equi_break_layer
        for blk in self.blocks[:half]:
            if hasattr(blk, "forward_packed"):
                t = blk.forward_packed(t)
            else:
                t = OF.pack_five(blk(OF.unpack_five(t)))
        B, N, D = t.shape
        if self.invariant:
            inv = self.invariantization.forward_packed(t)
            t = OF.LinearFn.apply(_rows(inv), self.invariant_proj.weight, self.invariant_proj.bias, False, True)
            t = t.view(B, N, D)
        else:
            t = OF.BridgeFn.apply(_rows(t)).view(B, N, D)
        taken = []
        for i, blk in enumerate(self.blocks[half:], start=half):
            t = blk(t)
            if i in take:
                taken.append(t)
        return t, taken

    def _final_norm(self, t: Tensor) -> Tensor:
        B, N, D = t.shape
        y = OF.LayerNormFn.apply(_rows(t), self.norm.weight, self.norm.bias, self.norm.eps, False, False)
        return y.view(B, N, D)

    def _output(self, x: Tensor, masks):
        x_norm = self._final_norm(x)
        R = self.num_register_tokens
        return {"x_norm_clstoken": x_norm[:, 0], "x_norm_regtokens": x_norm[:, 1:R + 1],
                "x_norm_patchtokens": x_norm[:, R + 1:], "x_prenorm": x, "masks": masks}

    concat_crops = True     # crop lists: run every per-token kernel once over the concatenated rows of all crops

    def _bridge_rows(self, rows: Tensor) -> Tensor:
        """octic half -> dense half on [T, D] rows: invariantisation + projection, or the 5 -> 8 tuple permutation"""
        if self.invariant:
            inv = self.invariantization.forward_packed(rows)
            return OF.LinearFn.apply(_rows(inv), self.invariant_proj.weight, self.invariant_proj.bias, False, True)
        return OF.BridgeFn.apply(rows)

    def forward_features_list(self, x_list, masks_list):
        """reference :138-168.  The reference hands the crop list to its NestedTensorBlock*s, which (with xformers)
        concatenate the crops along the token axis under a block-diagonal attention mask (dinov2/layers/block.py:
        212-248).  Here: the token rows of all crops are concatenated ONCE, every per-token kernel (LayerNorm, GEMMs,
        D8-GELU, residual epilogues, bridge) runs once over all rows and attention runs once per crop resolution on its
        row slice (`segs`); stochastic depth follows the reference's list rule with one draw per crop batch.  Models
        whose blocks cannot take segments (non-fused children, crop lengths outside the tcgen05 attention envelope)
        run crop by crop."""
        ts = [self.prepare_tokens_packed(x, m) for x, m in zip(x_list, masks_list)]
        half = self.octic_equi_break_layer
        octic, dense = self.blocks[:half], self.blocks[half:]
        nested = (all(isinstance(b, NestedTensorBlockD8) for b in octic)
                  and all(isinstance(b, NestedTensorBlock) for b in dense))
        segs = tuple((t.shape[0], t.shape[1]) for t in ts)
        if nested and self.concat_crops and len(ts) > 1 and all(b.supports_segments(segs) for b in octic):
            D = ts[0].shape[-1]
            t = torch.cat([x.reshape(-1, D) for x in ts]).unsqueeze(0)
            for blk in octic:
                t = blk.forward_packed(t, segs)
            t = self._bridge_rows(t[0]).view(1, -1, D)
            for blk in dense:
                t = _DenseBlockBase.forward(blk, t, True, segs)
            out, r0 = [], 0
            for b, n in segs:
                out.append(t[0, r0:r0 + b * n].view(b, n, D))
                r0 += b * n
        elif nested:
            for blk in octic:
                ts = [blk.forward_packed(t) for t in ts]
            out = [self._bridge_rows(_rows(t)).view(t.shape) for t in ts]
            for blk in dense:
                out = blk(out)
        else:
            out = [self._backbone_packed(t)[0] for t in ts]
        return [self._output(x, m) for x, m in zip(out, masks_list)]

    def forward_features(self, x, masks=None):
        """reference :170-198"""
        if isinstance(x, list):
            return self.forward_features_list(x, masks if masks is not None else [None] * len(x))
        t, _ = self._backbone_packed(self.prepare_tokens_packed(x, masks))
        return self._output(t, masks)

    def _get_intermediate_layers_not_chunked(self, x, n=1):
        """reference :200-227 (only dense-half blocks can be taken)"""
        total = len(self.blocks)
        blocks_to_take = range(total - n, total) if isinstance(n, int) else n
        assert all(i > self.octic_equi_break_layer for i in blocks_to_take), \
            f"All block indices must be > half depth, got {blocks_to_take}"
        _, output = self._backbone_packed(self.prepare_tokens_packed(x), take=tuple(blocks_to_take))
        assert len(output) == len(blocks_to_take), f"only {len(output)} / {len(blocks_to_take)} blocks found"
        return output

    def get_intermediate_layers(self, x: Tensor, n: Union[int, Sequence] = 1, reshape: bool = False,
                                return_class_token: bool = False, norm=True):
        """reference :229-253"""
        if self.chunked_blocks:
            raise NotImplementedError("Chunked blocks not supported yet")
        outputs = self._get_intermediate_layers_not_chunked(x, n)
        if norm:
            outputs = [self._final_norm(out) for out in outputs]
        class_tokens = [out[:, 0] for out in outputs]
        outputs = [out[:, 1 + self.num_register_tokens:] for out in outputs]
        if reshape:
            B, _, w, h = x.shape
            p = self.patch_embed.patch_size[0]
            outputs = [out.reshape(B, w // p, h // p, -1).permute(0, 3, 1, 2).contiguous() for out in outputs]
        if return_class_token:
            return tuple(zip(outputs, class_tokens))
        return tuple(outputs)

    @OF.opaque_to_compile
    def forward(self, *args, is_training=False, **kwargs):
        """reference :255-260"""
        ret = self.forward_features(*args, **kwargs)
        if is_training:
            return ret
        return self.head(ret["x_norm_clstoken"])

    @torch.jit.ignore
    def no_weight_decay(self):
        base_names = ['pos_embed.0', 'pos_embed.1', 'pos_embed.2', 'pos_embed.3', 'pos_embed.4', 'pos_embed.5',
                      'cls_token.0']
        return set(base_names + [f'_orig_mod.{name}' for name in base_names])


# ----------------------------------------------------------------------------------------------------------------
# factories (reference dinov2_models.py:269-329)
# ----------------------------------------------------------------------------------------------------------------
def _dino(patch_size, embed_dim, depth, num_heads, invariant, num_register_tokens, **kwargs):
    return OcticDinoVisionTransformer(
        patch_size=patch_size, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=4, invariant=invariant,
        standard_block_layers=partial(NestedTensorBlock, attn_class=MemEffAttention, init_values=1.0e-05),
        octic_block_layers=partial(NestedTensorBlockD8, init_values=1.0e-05),
        num_register_tokens=num_register_tokens, **kwargs)


@register_model
def hybrid_dinov2_vit_large_patch16(patch_size=16, num_register_tokens=0, **kwargs):
    return _dino(patch_size, 1024, 24, 16, False, num_register_tokens, **kwargs)


@register_model
def hybrid_dinov2_vit_huge_patch16(patch_size=16, num_register_tokens=0, **kwargs):
    return _dino(patch_size, 1280, 32, 16, False, num_register_tokens, **kwargs)


@register_model
def d8_inv_early_dinov2_vit_large_patch16(patch_size=16, num_register_tokens=0, **kwargs):
    return _dino(patch_size, 1024, 24, 16, True, num_register_tokens, **kwargs)


@register_model
def d8_inv_early_dinov2_vit_huge_patch16(patch_size=16, num_register_tokens=0, **kwargs):
    return _dino(patch_size, 1280, 32, 16, True, num_register_tokens, **kwargs)
