"""CPU oracle for the octic ViT block hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of the reference algorithm for the hot path of
davnords/octic-vits.  It is imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs, and only as the checker or the timed CPU baseline -- never by the product path
(octic_vits_b200/), which has no CPU fallback.

Parity status: PINNED.  tests/golden/*.pt hold inputs/outputs produced by running the reference's own modules
(imported from /root/reference with a timm stand-in, see tools/make_golden.py); tests/test_oracle_golden.py checks
every function below against them, plus the reference's own known-answer vector for GeluD8.

Conventions follow the reference: an octic feature is the 5-tuple (A1, A2, B1, B2, E) with shapes [B,N,C] x4 and
[B,N,2,2C]; "8-tuple" = (A1, A2, B1, B2, E11, E21, E12, E22), each [B,N,C].  Every function cites the reference
file:line it restates (paths relative to the reference repository root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Five = Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]

RT2_4 = math.sqrt(2.0) / 4.0
IRREPS = ("A1", "A2", "B1", "B2")


# ------------------------------------------------------------------------------------------------------------
# tuple plumbing  (octic_vits/d8_utils.py:358-385)
# ------------------------------------------------------------------------------------------------------------
def five_to_eight(xs: Five) -> List[Tensor]:
    e = xs[4]
    c = e.shape[-1] // 2
    return [xs[0], xs[1], xs[2], xs[3], e[..., 0, :c], e[..., 1, :c], e[..., 0, c:], e[..., 1, c:]]


def eight_to_five(ys: Sequence[Tensor]) -> Five:
    row0 = torch.cat((ys[4], ys[6]), dim=-1)
    row1 = torch.cat((ys[5], ys[7]), dim=-1)
    return (ys[0], ys[1], ys[2], ys[3], torch.stack((row0, row1), dim=-2))


def pack_rows(xs: Five) -> Tensor:
    """5-tuple -> packed rows [..., 8C] = [A1 | A2 | B1 | B2 | E row 0 | E row 1]  (layout of include/octic_b200.h)."""
    e = xs[4]
    return torch.cat((xs[0], xs[1], xs[2], xs[3], e[..., 0, :], e[..., 1, :]), dim=-1)


def unpack_rows(x: Tensor) -> Five:
    c = x.shape[-1] // 8
    a1, a2, b1, b2 = (x[..., i * c:(i + 1) * c] for i in range(4))
    e = torch.stack((x[..., 4 * c:6 * c], x[..., 6 * c:8 * c]), dim=-2)
    return (a1, a2, b1, b2, e)


# ------------------------------------------------------------------------------------------------------------
# D8 Fourier transforms and GELU  (octic_vits/d8_utils.py:276-356, octic_vits/d8_layers.py:98-102)
# ------------------------------------------------------------------------------------------------------------
# rows = outputs; the dense (no-FFT) form of d8_utils.py:305-315.  R2I is its transpose (the matrix is orthogonal
# after the 1/sqrt(8) scale), d8_utils.py:346-356.
_I2R_SIGNS = torch.tensor([
    [1, 1, 1, 1, 1, 1, 1, -1],
    [1, 1, -1, -1, 1, -1, -1, -1],
    [1, 1, 1, 1, -1, -1, -1, 1],
    [1, 1, -1, -1, -1, 1, 1, 1],
    [1, -1, 1, -1, -1, 1, -1, -1],
    [1, -1, -1, 1, -1, -1, 1, -1],
    [1, -1, 1, -1, 1, -1, 1, 1],
    [1, -1, -1, 1, 1, 1, -1, 1],
], dtype=torch.float64)


def isotypic_to_regular(xs8: Sequence[Tensor]) -> List[Tensor]:
    m = _I2R_SIGNS.to(dtype=xs8[0].dtype, device=xs8[0].device) * RT2_4
    x = torch.stack(list(xs8), dim=-1)
    y = x @ m.T
    return list(y.unbind(-1))


def regular_to_isotypic(xs8: Sequence[Tensor]) -> List[Tensor]:
    m = _I2R_SIGNS.to(dtype=xs8[0].dtype, device=xs8[0].device) * RT2_4
    x = torch.stack(list(xs8), dim=-1)
    y = x @ m
    return list(y.unbind(-1))


def gelu_d8_eight(xs8: Sequence[Tensor]) -> List[Tensor]:
    """GeluD8.forward on the 8-tuple (d8_layers.py:98-102): R2I(gelu(I2R(x))), exact erf GELU."""
    return regular_to_isotypic([F.gelu(v) for v in isotypic_to_regular(xs8)])


def gelu_d8(xs: Five) -> Five:
    """TritonGeluD8.forward on the 5-tuple (d8_gelu.py:456-482), via the index map of d8_gelu.py:517-541."""
    return eight_to_five(gelu_d8_eight(five_to_eight(xs)))


# ------------------------------------------------------------------------------------------------------------
# LinearD8 / AffineD8 / LayerNormD8 / LayerScaleD8  (octic_vits/d8_layers.py:104-212)
# ------------------------------------------------------------------------------------------------------------
def linear_d8(xs: Five, w: Dict[str, Tensor], prefix: str) -> Five:
    """LinearD8.forward (d8_layers.py:124-127).  `w` is a reference-style state dict; keys
    `{prefix}lin_{A1,A2,B1,B2,E}.weight`, optional `{prefix}lin_A1.bias`."""
    outs = []
    for i, name in enumerate(IRREPS):
        b = w.get(f"{prefix}lin_{name}.bias") if name == "A1" else None
        outs.append(F.linear(xs[i], w[f"{prefix}lin_{name}.weight"], b))
    outs.append(F.linear(xs[4], w[f"{prefix}lin_E.weight"]))
    return tuple(outs)


def affine_d8(xs: Five, w: Dict[str, Tensor], prefix: str) -> Five:
    """AffineD8.forward (d8_layers.py:147-158); LayerScaleD8.forward (:205-212) is the beta-less case."""
    beta = w.get(f"{prefix}beta")
    y0 = w[f"{prefix}alpha_A1"] * xs[0]
    if beta is not None:
        y0 = y0 + beta
    return (y0, w[f"{prefix}alpha_A2"] * xs[1], w[f"{prefix}alpha_B1"] * xs[2], w[f"{prefix}alpha_B2"] * xs[3],
            w[f"{prefix}alpha_E"] * xs[4])


def layernorm_d8(xs: Five, w: Dict[str, Tensor], prefix: str, eps: float = 1e-5) -> Five:
    """LayerNormD8.forward (d8_layers.py:166-186): six means, one shared std
    std = sqrt(2)/4 * sqrt(sum_i var(x_i) + mean_r var(E[r]) + eps)."""
    def var(t):
        return t.var(dim=-1, unbiased=False, keepdim=True)
    s = var(xs[0]) + var(xs[1]) + var(xs[2]) + var(xs[3]) + var(xs[4]).mean(dim=-2) + eps
    std = RT2_4 * torch.sqrt(s)
    normed = tuple((xs[i] - xs[i].mean(dim=-1, keepdim=True)) / std for i in range(4)) + (
        (xs[4] - xs[4].mean(dim=-1, keepdim=True)) / std.unsqueeze(-1),)
    return affine_d8(normed, w, f"{prefix}scaling.")


# ------------------------------------------------------------------------------------------------------------
# AttentionD8 / MlpD8 / blocks  (octic_vits/d8_layers.py:215-247, 590-776)
# ------------------------------------------------------------------------------------------------------------
def attention_heads_d8(qkvs: Five, num_heads: int) -> Tuple[Tensor, Tensor, Tensor]:
    """q, k, v [B,H,N,hd] from the LinearD8(dim, 3*dim) output (d8_layers.py:632-643).  Per head the vector is
    [A1 (c_h) | A2 | B1 | B2 | E row 0 (2 c_h) | E row 1 (2 c_h)]; out-features are ordered [3][H][c_h]."""
    B, N, C3 = qkvs[0].shape
    C, H = C3 // 3, num_heads
    ch = C // H
    parts = [t.reshape(B, N, 3, H, ch) for t in qkvs[:4]]
    e = qkvs[4].reshape(B, N, 2, 3, H, 2 * ch)             # [B,N,row,3,H,2ch]
    parts.append(e[:, :, 0])                               # row 0 -> [B,N,3,H,2ch]
    parts.append(e[:, :, 1])
    full = torch.cat(parts, dim=-1)                        # [B,N,3,H,hd]
    full = full.permute(2, 0, 3, 1, 4)                     # [3,B,H,N,hd]
    return full[0], full[1], full[2]


def attention_unpack_d8(x: Tensor) -> Five:
    """[B,H,N,hd] -> 5-tuple (d8_layers.py:650-656)."""
    B, H, N, hd = x.shape
    ch = hd // 8
    C = H * ch
    outs = []
    for i in range(4):
        outs.append(x[..., i * ch:(i + 1) * ch].transpose(1, 2).reshape(B, N, C))
    rows = []
    for r in range(2):
        seg = x[..., 4 * ch + r * 2 * ch: 4 * ch + (r + 1) * 2 * ch]     # [B,H,N,2ch]
        rows.append(seg.transpose(1, 2).reshape(B, N, 2 * C))
    outs.append(torch.stack(rows, dim=-2))
    return tuple(outs)


def sdpa(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """softmax(q k^T / sqrt(hd)) v -- the SDPA default scale; AttentionD8.scale is unused (d8_layers.py:613,645)."""
    s = (q @ k.transpose(-2, -1)) / math.sqrt(q.shape[-1])
    return torch.softmax(s, dim=-1) @ v


def attention_d8(xs: Five, w: Dict[str, Tensor], prefix: str, num_heads: int) -> Five:
    """AttentionD8.forward (d8_layers.py:623-660), dropout p = 0."""
    q, k, v = attention_heads_d8(linear_d8(xs, w, f"{prefix}qkv."), num_heads)
    return linear_d8(attention_unpack_d8(sdpa(q, k, v)), w, f"{prefix}proj.")


def mlp_d8(xs: Five, w: Dict[str, Tensor], prefix: str) -> Five:
    """MlpD8.forward (d8_layers.py:240-247) with DropoutD8(p=0) and norm=Identity."""
    return linear_d8(gelu_d8(linear_d8(xs, w, f"{prefix}fc1.")), w, f"{prefix}fc2.")


def _add(xs: Five, ys: Five) -> Five:
    return tuple(a + b for a, b in zip(xs, ys))


def _scale_rows(xs: Five, s: Optional[Tensor]) -> Five:
    """drop_path_d8 (d8_layers.py:249-271) with an explicit per-sample factor s [B] = mask / keep_prob."""
    if s is None:
        return xs
    return tuple(t * s.reshape(-1, *([1] * (t.dim() - 1))) for t in xs)


def block_d8(xs: Five, w: Dict[str, Tensor], prefix: str, num_heads: int, style: str = "deit",
             drop_scale1: Optional[Tensor] = None, drop_scale2: Optional[Tensor] = None) -> Five:
    """Layer_scale_init_BlockD8.forward (d8_layers.py:703-707; style='deit', layer scale = gamma_1/2) or
    BlockD8.forward (d8_layers.py:759-776; style='dinov2', layer scale = ls1/ls2, absent when init_values is falsy)."""
    ls1, ls2 = ("gamma_1.", "gamma_2.") if style == "deit" else ("ls1.", "ls2.")
    y = attention_d8(layernorm_d8(xs, w, f"{prefix}norm1."), w, f"{prefix}attn.", num_heads)
    if f"{prefix}{ls1}alpha_A1" in w:
        y = affine_d8(y, w, f"{prefix}{ls1}")
    xs = _add(xs, _scale_rows(y, drop_scale1))
    y = mlp_d8(layernorm_d8(xs, w, f"{prefix}norm2."), w, f"{prefix}mlp.")
    if f"{prefix}{ls2}alpha_A1" in w:
        y = affine_d8(y, w, f"{prefix}{ls2}")
    return _add(xs, _scale_rows(y, drop_scale2))


# ------------------------------------------------------------------------------------------------------------
# invariantisation and bridge  (octic_vits/d8_invariantization.py:49-64, octic_vits/model.py:196-200)
# ------------------------------------------------------------------------------------------------------------
def power_spectrum(xs: Five) -> Tensor:
    return torch.cat((xs[0], xs[1].abs(), xs[2].abs(), xs[3].abs(), xs[4].norm(dim=-2)), dim=-1)


def hybrid_bridge(xs: Five) -> Tensor:
    return torch.cat(five_to_eight(xs), dim=-1)


# ------------------------------------------------------------------------------------------------------------
# dense half  (deit/vit.py:14-56, 90-134; timm Block has the same maths with ls1/ls2.gamma)
# ------------------------------------------------------------------------------------------------------------
def dense_attention(x: Tensor, w: Dict[str, Tensor], prefix: str, num_heads: int) -> Tensor:
    B, N, D = x.shape
    qkv = F.linear(x, w[f"{prefix}qkv.weight"], w.get(f"{prefix}qkv.bias"))
    qkv = qkv.reshape(B, N, 3, num_heads, D // num_heads).permute(2, 0, 3, 1, 4)
    o = sdpa(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(B, N, D)
    return F.linear(o, w[f"{prefix}proj.weight"], w.get(f"{prefix}proj.bias"))


def dense_block(x: Tensor, w: Dict[str, Tensor], prefix: str, num_heads: int, eps: float = 1e-6,
                drop_scale1: Optional[Tensor] = None, drop_scale2: Optional[Tensor] = None) -> Tensor:
    """Layer_scale_init_Block.forward (deit/vit.py:131-134): x += dp(gamma_1 * attn(norm1 x)); x += dp(gamma_2 * mlp(norm2 x))."""
    D = x.shape[-1]
    g1 = w.get(f"{prefix}gamma_1", w.get(f"{prefix}ls1.gamma"))
    g2 = w.get(f"{prefix}gamma_2", w.get(f"{prefix}ls2.gamma"))
    y = dense_attention(F.layer_norm(x, (D,), w[f"{prefix}norm1.weight"], w[f"{prefix}norm1.bias"], eps), w,
                        f"{prefix}attn.", num_heads)
    if g1 is not None:
        y = g1 * y
    if drop_scale1 is not None:
        y = y * drop_scale1.reshape(-1, 1, 1)
    x = x + y
    h = F.layer_norm(x, (D,), w[f"{prefix}norm2.weight"], w[f"{prefix}norm2.bias"], eps)
    h = F.linear(F.gelu(F.linear(h, w[f"{prefix}mlp.fc1.weight"], w[f"{prefix}mlp.fc1.bias"])),
                 w[f"{prefix}mlp.fc2.weight"], w[f"{prefix}mlp.fc2.bias"])
    if g2 is not None:
        h = g2 * h
    if drop_scale2 is not None:
        h = h * drop_scale2.reshape(-1, 1, 1)
    return x + h


# ------------------------------------------------------------------------------------------------------------
# front end: lifting patch embedding and symmetric positional embedding
# (octic_vits/d8_layers.py:284-486, octic_vits/d8_utils.py:388-451, octic_vits/model.py:170-181)
# ------------------------------------------------------------------------------------------------------------
def _unfold_quadrant(w: Tensor, sign_rot: float, sign_flip: float, dims: Tuple[int, int]) -> Tensor:
    """Tile a quadrant `w` into the full square: [[w, s*rot3(w)], [s*rot1(w), rot2(w)]] (blocks along dims), then
    add sign_flip * (mirror along dims[1]).  Shared by expand_weight (d8_layers.py:334-373) and
    isotypic_dim_interpolation (d8_utils.py:401-431)."""
    d0, d1 = dims
    left = torch.cat((w, sign_rot * w.rot90(1, dims)), dim=d0)
    right = torch.cat((sign_rot * w.rot90(3, dims), w.rot90(2, dims)), dim=d0)
    full = torch.cat((left, right), dim=d1)
    return full + sign_flip * full.flip(d1)


_SIGNS = {"A1": (1.0, 1.0), "A2": (1.0, -1.0), "B1": (-1.0, 1.0), "B2": (-1.0, -1.0)}


def expand_lift_weight(weight: Tensor, irrep: str) -> Tensor:
    """LiftIrrepD8Conv2d.expand_weight (d8_layers.py:329-373): half-size filters [Co,Ci,p/2,p/2] -> [Co,Ci,p,p]."""
    if irrep == "E":
        w = 0.5 * weight
        w2 = torch.cat((w, w.flip(-2)), dim=-2)
        return torch.cat((w2, -w2.flip(-1)), dim=-1)
    sr, sf = _SIGNS[irrep]
    return _unfold_quadrant(RT2_4 * weight, sr, sf, (-2, -1))


def patch_embed_d8(img: Tensor, w: Dict[str, Tensor], prefix: str, patch: int) -> Five:
    """PatchEmbedD8.forward (d8_layers.py:452-486): 8 stride-p convolutions with D8-symmetrised filters; the E
    convolutions run twice, the second time with the filter rotated by 90 degrees (d8_layers.py:377-381)."""
    outs = []
    for name in IRREPS:
        k = expand_lift_weight(w[f"{prefix}lift8.conv_{name}.weight"], name)
        b = w.get(f"{prefix}lift8.conv_{name}.bias")
        outs.append(F.conv2d(img, k, b, stride=patch))
    for side in ("E_left", "E_right"):
        k = expand_lift_weight(w[f"{prefix}lift8.conv_{side}.weight"], "E")
        outs.append(F.conv2d(img, k, None, stride=patch))
        outs.append(F.conv2d(img, k.rot90(1, (-2, -1)), None, stride=patch))
    flat = [t.flatten(2).transpose(1, 2) for t in outs]     # BCHW -> BNC
    return eight_to_five(flat)


def isotypic_to_patch(xs: Five, w: Dict[str, Tensor], prefix: str, patch_side: int, out_channels: int = 3,
                      reshape_to_image: bool = False) -> Tensor:
    """IsotypicToPatchD8.forward (d8_layers.py:520-588): LinearD8(dim, 2 p^2 c) -> per irrep a quadrant
    [B, L, p/2, p/2, c] (scaled 1/4) unfolded to the full patch with the irrep's symmetry; the E irrep contributes its
    first two components (sqrt(2) x4 as is, sqrt(2) x5 rotated by 90 degrees), x6/x7 are not used."""
    ys = five_to_eight(linear_d8(xs, w, f"{prefix}lin8."))
    B, L, _ = ys[0].shape
    h = patch_side // 2
    q = [0.25 * y.reshape(B, L, h, h, out_channels) for y in ys]
    out = sum(_unfold_quadrant(q[i], *_SIGNS[name], (2, 3)) for i, name in enumerate(IRREPS))
    for i, turns in ((4, 0), (5, 1)):
        x = math.sqrt(2.0) * q[i]
        col = torch.cat((x, x.flip(2)), dim=2)
        full = torch.cat((col, -col.flip(3)), dim=3)
        out = out + (full.rot90(turns, (2, 3)) if turns else full)
    if reshape_to_image:
        H = W = int(math.isqrt(L))
        out = out.reshape(B, H, W, patch_side, patch_side, out_channels)
        return out.permute(0, 5, 1, 3, 2, 4).reshape(B, out_channels, H * patch_side, W * patch_side)
    return out.reshape(B, L, patch_side ** 2 * out_channels)


def unfold_pos_embed(ps: Sequence[Tensor]) -> Five:
    """isotypic_dim_interpolation(dim=0) + convert_8tuple_to_5tuple (d8_utils.py:388-451, model.py:174):
    six [h/2, w/2, C] parameters -> 5-tuple of [h, w, C] x4 and [h, w, 2, 2C]."""
    outs = [_unfold_quadrant(ps[i], *_SIGNS[name], (0, 1)) for i, name in enumerate(IRREPS)]
    for p in (ps[4], ps[5]):
        col = torch.cat((p, p.flip(0)), dim=0)
        full = torch.cat((col, -col.flip(1)), dim=1)
        outs.append(full)
        outs.append(full.rot90(1, (0, 1)))
    return eight_to_five(outs)


def embed_tokens(img: Tensor, w: Dict[str, Tensor], patch: int) -> Five:
    """patch embed + pos embed + cls token (model.py:172-181), global_pool=False, no registers."""
    xs = patch_embed_d8(img, w, "patch_embed.", patch)
    pos = unfold_pos_embed([w[f"pos_embed.{i}"] for i in range(6)])
    xs = tuple(x + p.flatten(0, 1) for x, p in zip(xs, pos))
    B = img.shape[0]
    out = []
    for i in range(5):
        cls = w[f"cls_token.{i}"]
        out.append(torch.cat((cls.expand(B, *cls.shape[1:]), xs[i]), dim=1))
    return tuple(out)


# ------------------------------------------------------------------------------------------------------------
# whole model  (octic_vits/model.py:170-227)
# ------------------------------------------------------------------------------------------------------------
def octic_vit_forward(img: Tensor, w: Dict[str, Tensor], *, patch: int, depth: int, num_heads: int,
                      invariant: bool = False, break_layer: Optional[int] = None, style: str = "deit",
                      return_trunk: bool = False):
    """OcticVisionTransformer.forward (model.py:170-227), eval mode.  `w` is the reference state dict."""
    k = depth // 2 if break_layer is None else break_layer
    xs = embed_tokens(img, w, patch)
    for i in range(k):
        xs = block_d8(xs, w, f"blocks.{i}.", num_heads, style)
    if return_trunk:
        return xs
    if invariant:
        x = F.linear(power_spectrum(xs), w["invariant_proj.weight"], w["invariant_proj.bias"])
    else:
        x = hybrid_bridge(xs)
    for i in range(k, depth):
        x = dense_block(x, w, f"blocks.{i}.", num_heads)
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), w["norm.weight"], w["norm.bias"], 1e-6)[:, 0]
    if "head.weight" in w:
        x = F.linear(x, w["head.weight"], w["head.bias"])
    return x


# ------------------------------------------------------------------------------------------------------------
# DINOv2 backbone  (octic_vits/dinov2_models.py:40-260)
# ------------------------------------------------------------------------------------------------------------
def dino_prepare_tokens(img: Tensor, w: Dict[str, Tensor], patch: int, masks: Optional[Tensor] = None) -> Five:
    """OcticDinoVisionTransformer.prepare_tokens_with_masks (dinov2_models.py:113-136), native resolution:
    patch embed -> iBOT mask-token substitution on the 8-tuple -> + unfolded pos-embed -> cls token (no pos-embed,
    'deviating from DINOv2') -> register tokens between cls and patches."""
    xs8 = five_to_eight(patch_embed_d8(img, w, "patch_embed.", patch))
    if masks is not None:
        xs8 = [torch.where(masks.unsqueeze(-1), w[f"mask_token.{i}"].to(x.dtype).unsqueeze(0), x)
               for i, x in enumerate(xs8)]
    pos8 = five_to_eight(unfold_pos_embed([w[f"pos_embed.{i}"] for i in range(6)]))
    h0, w0 = img.shape[-2] // patch, img.shape[-1] // patch
    if pos8[0].shape[0] * pos8[0].shape[1] != xs8[0].shape[1] or h0 != w0:
        # interpolate_spatial_tuple (d8_utils.py:453-499) as written; the reference never reaches it (its callers pass
        # the patch-size tuple), so this branch is UNPINNED by reference outputs
        st = torch.stack([p.float() for p in pos8], dim=0).permute(0, 3, 1, 2)
        st = F.interpolate(st, size=(h0, w0), mode="bicubic", antialias=False).permute(0, 2, 3, 1)
        pos8 = [st[i].to(xs8[0].dtype) for i in range(8)]
    xs8 = [x + p.flatten(0, 1) for x, p in zip(xs8, pos8)]
    B = img.shape[0]
    xs8 = [torch.cat((w[f"cls_token.{i}"].expand(B, -1, -1), x), dim=1) for i, x in enumerate(xs8)]
    if "register_tokens.0" in w:
        xs8 = [torch.cat((x[:, :1], w[f"register_tokens.{i}"].expand(B, -1, -1), x[:, 1:]), dim=1)
               for i, x in enumerate(xs8)]
    return eight_to_five(xs8)


def octic_dino_forward_features(img: Tensor, w: Dict[str, Tensor], *, patch: int, depth: int, num_heads: int,
                                invariant: bool = False, masks: Optional[Tensor] = None,
                                drop_scales: Optional[Sequence[Tuple[Optional[Tensor], Optional[Tensor]]]] = None,
                                take: Sequence[int] = ()):
    """forward_features (dinov2_models.py:170-198): octic BlockD8 x depth/2 (the subclass ignores
    octic_equi_break_layer), bridge or invariantisation, dense NestedTensorBlock x depth/2
    (dinov2/layers/block.py:43-114; same maths as dense_block with ls1/ls2.gamma), LayerNorm(eps=1e-6).
    `drop_scales[i]` = explicit per-sample residual factors of block i (stochastic depth with injected draws).
    `take` = dense block indices whose outputs are also returned (get_intermediate_layers, :200-227)."""
    R = w["register_tokens.0"].shape[1] if "register_tokens.0" in w else 0
    ds = drop_scales or [(None, None)] * depth
    xs = dino_prepare_tokens(img, w, patch, masks)
    for i in range(depth // 2):
        xs = block_d8(xs, w, f"blocks.{i}.", num_heads, "dinov2", ds[i][0], ds[i][1])
    if invariant:
        x = F.linear(power_spectrum(xs), w["invariant_proj.weight"], w["invariant_proj.bias"])
    else:
        x = hybrid_bridge(xs)
    taken = []
    for i in range(depth // 2, depth):
        x = dense_block(x, w, f"blocks.{i}.", num_heads, 1e-6, ds[i][0], ds[i][1])
        if i in take:
            taken.append(x)
    xn = F.layer_norm(x, (x.shape[-1],), w["norm.weight"], w["norm.bias"], 1e-6)
    out = {"x_norm_clstoken": xn[:, 0], "x_norm_regtokens": xn[:, 1:R + 1], "x_norm_patchtokens": xn[:, R + 1:],
           "x_prenorm": x, "masks": masks}
    return (out, taken) if take else out


# ------------------------------------------------------------------------------------------------------------
# group actions used by the equivariance tests  (octic_vits/d8_utils.py:76-274)
# ------------------------------------------------------------------------------------------------------------
GROUP = ("e", "r", "rr", "rrr", "m", "mr", "mrr", "mrrr")


def image_action(g: str, img: Tensor) -> Tensor:
    """d8_utils.py:76-94: r = rot90 on the last two dims, m = flip of the last dim applied after the rotations."""
    k = g.count("r")
    out = img.rot90(k, (-2, -1)) if k else img
    return out.flip(-1) if g.startswith("m") else out


def isotypic_action(g: str, xs8: Sequence[Tensor]) -> List[Tensor]:
    """d8_utils.py:179-260.  One application of r: (x4,x5)->(-x5,x4), (x6,x7)->(-x7,x6), B1,B2 -> -B1,-B2;
    m: A2,B2 -> -A2,-B2, (x4,x5)->(-x4,x5), (x6,x7)->(-x6,x7).  `g` = 'm'? followed by r's means: rotate first,
    then mirror (matches image_action)."""
    a1, a2, b1, b2, x4, x5, x6, x7 = xs8
    for _ in range(g.count("r")):
        b1, b2 = -b1, -b2
        x4, x5 = -x5, x4
        x6, x7 = -x7, x6
    if g.startswith("m"):
        a2, b2 = -a2, -b2
        x4, x6 = -x4, -x6
    return [a1, a2, b1, b2, x4, x5, x6, x7]


def token_action(g: str, xs: Five, has_cls: bool = True) -> Five:
    """spatial_and_isotypic_group_action (d8_utils.py:262-274) on a 5-tuple; a leading cls token, if present, only
    sees the isotypic action."""
    xs8 = five_to_eight(xs)
    out = []
    for t in xs8:
        cls, pat = (t[:, :1], t[:, 1:]) if has_cls else (t[:, :0], t)
        B, L, C = pat.shape
        s = int(math.isqrt(L))
        grid = pat.transpose(1, 2).reshape(B, C, s, s)
        grid = image_action(g, grid)
        out.append(torch.cat((cls, grid.flatten(2).transpose(1, 2)), dim=1))
    return eight_to_five(isotypic_action(g, out))
