"""OcticVisionTransformer with the reference's constructor and state-dict (octic_vits/model.py:25-234), running on the
packed-row sm_100a kernels.  The octic trunk keeps ONE fp32 [B, N, D] token matrix from the patch embedding to the
bridge; no 5-tuples are materialised inside the model.
"""
from __future__ import annotations

from functools import partial
from typing import Callable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as OF
from .layers import (SQRT2_OVER_2, Block, BlockD8, LayerNormD8, PatchEmbedD8, PowerSpectrumInvariant, TritonGeluD8,
                     _rows)


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def unfold_pos(p: torch.Tensor, kind: str) -> torch.Tensor:
    """One stored positional-embedding quadrant [h/2, w/2, C] -> the full grid [h, w, C] of irrep component `kind`
    (isotypic_dim_interpolation(dim=0), reference octic_vits/d8_utils.py:388-451); "Er" = the second row of the E pair."""
    if kind in ("E", "Er"):
        col = torch.cat((p, p.flip(0)), dim=0)
        full = torch.cat((col, -col.flip(1)), dim=1)
        return full.rot90(1, (0, 1)) if kind == "Er" else full
    s_rot = 1.0 if kind in ("A1", "A2") else -1.0
    s_flip = 1.0 if kind in ("A1", "B1") else -1.0
    left = torch.cat((p, s_rot * p.rot90(1, (0, 1))), dim=0)
    right = torch.cat((s_rot * p.rot90(3, (0, 1)), p.rot90(2, (0, 1))), dim=0)
    full = torch.cat((left, right), dim=1)
    return full + s_flip * full.flip(1)


_POS_MAPS: dict = {}


def pos_maps(h2: int, w2: int, device):
    """ops.SparseMap per component kind over the positions of an h2 x w2 quadrant (one-hot basis pushed through unfold_pos)."""
    key = (h2, w2, str(device))
    if key not in _POS_MAPS:
        n = h2 * w2
        basis = torch.eye(n).view(h2, w2, n)
        _POS_MAPS[key] = {k: OF.ops.build_sparse_map(unfold_pos(basis, k).reshape(4 * n, n), device)
                          for k in ("A1", "A2", "B1", "B2", "E", "Er")}
    return _POS_MAPS[key]


def unfold_pos_embed_packed(ps) -> torch.Tensor:
    """isotypic_dim_interpolation(dim=0) + convert_8tuple_to_5tuple (reference octic_vits/d8_utils.py:388-451,
    model.py:174) producing packed rows [h*w, D] directly: A1 | A2 | B1 | B2 | E row0 = (x4, x6) | E row1 = (x5, x7).
    On the GPU one sparse-map launch per column block (functional.PosEmbedFn); plain differentiable torch ops otherwise."""
    if ps[0].is_cuda:
        return OF.PosEmbedFn.apply(*ps, pos_maps(ps[0].shape[0], ps[0].shape[1], ps[0].device))
    packed = torch.cat((unfold_pos(ps[0], "A1"), unfold_pos(ps[1], "A2"), unfold_pos(ps[2], "B1"), unfold_pos(ps[3], "B2"),
                        unfold_pos(ps[4], "E"), unfold_pos(ps[5], "E"), unfold_pos(ps[4], "Er"), unfold_pos(ps[5], "Er")), dim=-1)
    return packed.flatten(0, 1)


class OcticVisionTransformer(nn.Module):
    """Same arguments as the reference (model.py:49-70).  `octic_block_layers` / `standard_block_layers` /
    `Patch_layer` default to this package's kernels-backed classes; any callable with the reference's block signature
    is accepted (octic blocks need `forward_packed`, otherwise they are fed 5-tuple views)."""

    def __init__(self, img_size: int = 224, patch_size: int = 16, in_chans: int = 3, num_classes: int = 1000,
                 embed_dim: int = 768, depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.,
                 qkv_bias: bool = False, drop_rate: float = 0., attn_drop_rate: float = 0., drop_path_rate: float = 0.,
                 octic_block_layers: Callable = BlockD8, standard_block_layers: Callable = Block,
                 Patch_layer: Callable = PatchEmbedD8, init_scale: float = 1e-4, num_register_tokens: int = 0,
                 global_pool: bool = False, invariant: bool = False, octic_equi_break_layer: Optional[int] = None,
                 **kwargs):
        super().__init__()
        assert embed_dim % 8 == 0, "embed_dim must be divisible by 8"
        self.dropout_rate = drop_rate
        self.global_pool = global_pool
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        if octic_equi_break_layer is None:
            assert depth % 2 == 0, "depth must be even"
            octic_equi_break_layer = depth // 2
        else:
            assert octic_equi_break_layer >= 0, "octic_equi_break_layer must be non-negative"
            assert octic_equi_break_layer < depth, "octic_equi_break_layer must be less than depth"
        self.octic_equi_break_layer = octic_equi_break_layer
        self.invariant = invariant
        self.num_register_tokens = num_register_tokens
        # optional callable applied (as a tensor hook) to the gradient that flows from the dense half into the octic
        # half: parallel.install_early_allreduce starts the gradient exchange of the dense half there
        self._bridge_grad_hook = None
        self._block_grad_hooks = {}      # block index -> hook on the gradient flowing into that block (more exchange buckets)
        if num_register_tokens > 0 and type(self) is OcticVisionTransformer:
            # the reference's base-class register path indexes range(8) over a 5-tuple and cannot run
            # (SURVEY.md Appendix A.5); only the DINOv2 subclass (dinov2_models.py) supports registers.
            raise NotImplementedError("register tokens are only supported by the DINOv2 wrapper in the reference")

        if self.invariant:
            self.invariantization = PowerSpectrumInvariant(embed_dim)
            self.invariant_proj = nn.Linear(self.invariantization.output_dim, embed_dim)

        self.patch_embed = Patch_layer(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        norm_layer = partial(nn.LayerNorm, eps=1e-6)

        if not global_pool:
            self.cls_token = nn.ParameterList(
                [nn.Parameter(torch.zeros(1, 1, embed_dim // 8), requires_grad=(i == 0)) for i in range(4)] +
                [nn.Parameter(torch.zeros(1, 1, 2, embed_dim // 4), requires_grad=False)])
        assert num_register_tokens >= 0
        if num_register_tokens > 0:     # reference model.py:106-110 (kept for the state-dict key order of subclasses)
            self.register_tokens = nn.ParameterList(
                [nn.Parameter(torch.zeros(1, num_register_tokens, embed_dim // 8), requires_grad=(i == 0))
                 for i in range(8)])
        self.pos_embed = nn.ParameterList([
            nn.Parameter(torch.empty(img_size // patch_size // 2, img_size // patch_size // 2, embed_dim // 8))
            for _ in range(6)])

        dpr = [drop_path_rate for _ in range(depth)]
        self.blocks = nn.ModuleList([
            octic_block_layers(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                               attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=LayerNormD8,
                               act_layer=TritonGeluD8, init_values=init_scale)
            if i < self.octic_equi_break_layer else
            standard_block_layers(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                  attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer,
                                  act_layer=nn.GELU, init_values=init_scale)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()

        std = 8 * .02
        for p in self.pos_embed:
            trunc_normal_(p, std=SQRT2_OVER_2 * std)
        if not global_pool:
            for p in self.cls_token:
                if p.requires_grad:
                    trunc_normal_(p, std=std)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.weight, 1.0)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    # ------------------------------------------------------------------------------------------------------------
    def embed_tokens_packed(self, x: torch.Tensor) -> torch.Tensor:
        """patch embed + symmetric pos-embed + cls token (reference model.py:172-181) -> fp32 packed [B, N, D]."""
        B = x.shape[0]
        lead = 0 if self.global_pool else 1
        pos = unfold_pos_embed_packed(self.pos_embed)                                  # [np, D]
        if pos.shape[0] != self.patch_embed.num_patches:
            raise NotImplementedError("positional-embedding interpolation (the reference path is broken, SURVEY App. B)")
        if lead:
            c = self.cls_token
            cls = torch.cat((c[0], c[1], c[2], c[3], c[4][:, :, 0, :], c[4][:, :, 1, :]), dim=-1)   # [1, 1, D]
            rows = torch.cat((cls[0], pos), dim=0)
        else:
            rows = pos
        tokens = rows.unsqueeze(0).expand(B, -1, -1).contiguous()
        if hasattr(self.patch_embed, "embed_into"):
            return self.patch_embed.embed_into(x, tokens, lead)
        xs = OF.pack_five(self.patch_embed(x))
        return torch.cat((tokens[:, :lead], tokens[:, lead:] + xs), dim=1)

    def forward_trunk_packed(self, x: torch.Tensor) -> torch.Tensor:
        """octic half (reference model.py:172-194) -> fp32 packed [B, N, D]"""
        OF.require_cuda(x)
        t = self.embed_tokens_packed(x)
        for i, blk in enumerate(self.blocks[:self.octic_equi_break_layer]):
            if i in self._block_grad_hooks and t.requires_grad:
                t.register_hook(self._block_grad_hooks[i])
            if hasattr(blk, "forward_packed"):
                t = blk.forward_packed(t)
            else:
                t = OF.pack_five(blk(OF.unpack_five(t)))
        return t

    @OF.opaque_to_compile
    def forward_features(self, x):
        OF.check_autocast()
        t = self.forward_trunk_packed(x)
        B, N, D = t.shape
        if self.invariant:
            inv = self.invariantization.forward_packed(t)
            t = OF.LinearFn.apply(_rows(inv), self.invariant_proj.weight, self.invariant_proj.bias, False, True)
            t = t.view(B, N, D)
        else:
            t = OF.BridgeFn.apply(_rows(t)).view(B, N, D)
        if self._bridge_grad_hook is not None and t.requires_grad:
            t.register_hook(self._bridge_grad_hook)
        for i, blk in enumerate(self.blocks[self.octic_equi_break_layer:], start=self.octic_equi_break_layer):
            if i in self._block_grad_hooks and i != self.octic_equi_break_layer and t.requires_grad:
                t.register_hook(self._block_grad_hooks[i])
            t = blk(t)
        if self.global_pool:
            t = OF.LayerNormFn.apply(_rows(t), self.norm.weight, self.norm.bias, self.norm.eps, False, False)
            return t.view(B, N, D).mean(dim=1)
        # LayerNorm is per token, so norm(x)[:, 0] == norm(x[:, 0]) (reference model.py:204-211)
        cls = t[:, 0]
        return OF.LayerNormFn.apply(cls, self.norm.weight, self.norm.bias, self.norm.eps, False, False)

    @OF.opaque_to_compile
    def forward(self, x):
        x = self.forward_features(x)
        if self.dropout_rate:
            x = F.dropout(x, p=float(self.dropout_rate), training=self.training)
        if isinstance(self.head, nn.Linear):
            x = OF.LinearFn.apply(x.to(torch.bfloat16), self.head.weight, self.head.bias, False, True)
        return x

    @torch.jit.ignore
    def no_weight_decay(self):
        base_names = ['pos_embed.0', 'pos_embed.1', 'pos_embed.2', 'pos_embed.3', 'pos_embed.4', 'pos_embed.5',
                      'cls_token.0']
        return set(base_names + [f'_orig_mod.{name}' for name in base_names])
