"""Generate tests/golden/*.pt by running the REFERENCE implementation (imported from /root/reference) on CPU.

Run in the build container only (`python tools/make_golden.py`); /root/reference does not exist on the GPU box, which
is why the vectors are committed.  Two stand-ins are needed to import the reference here:
tools/timm_shim (timm is not installed) and a CPU activation in place of the CUDA-only TritonGeluD8, built from the
reference's own GeluD8 + tuple converters exactly the way octic_vits/d8_gelu.py:517-541 maps between them.

Every fixture stores: the reference state dict, the inputs, the reference outputs and (where useful) reference
autograd gradients, all fp32.
"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
REF = Path(os.environ.get("OCTIC_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT / "tools" / "timm_shim"))
sys.path.insert(0, str(REF))

from octic_vits import d8_layers, d8_utils  # noqa: E402
from octic_vits.d8_invariantization import PowerSpectrumInvariant  # noqa: E402
from octic_vits.model import OcticVisionTransformer  # noqa: E402
from deit.vit import Layer_scale_init_Block  # noqa: E402


class CpuGeluD8(torch.nn.Module):
    """5-tuple wrapper around the reference GeluD8 (d8_layers.py:98-102)."""

    def __init__(self):
        super().__init__()
        self.inner = d8_layers.GeluD8()

    def forward(self, xs):
        return d8_utils.convert_8tuple_to_5tuple(self.inner(d8_utils.convert_5tuple_to_8tuple(xs)))


def swap_gelu(module):
    for name, child in module.named_children():
        if isinstance(child, d8_layers.TritonGeluD8):
            setattr(module, name, CpuGeluD8())
        else:
            swap_gelu(child)
    return module


def randomize(module, gen, std=0.3):
    """O(1) parameters everywhere so that no branch is ~identity (SURVEY.md section 4 caveat)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if not p.requires_grad:
                continue
            if "alpha" in name or name.endswith("gamma_1") or name.endswith("gamma_2") or ".gamma" in name:
                p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=gen))
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=gen))
            else:
                p.copy_(std * torch.randn(p.shape, generator=gen))


def rand5(B, N, C, gen):
    return tuple(torch.randn(B, N, C, generator=gen) for _ in range(4)) + (torch.randn(B, N, 2, 2 * C, generator=gen),)


def with_grads(module, xs, gen):
    xs = tuple(x.clone().requires_grad_(True) for x in xs)
    out = module(xs)
    gs = tuple(torch.randn(o.shape, generator=gen) for o in out)
    loss = sum((o * g).sum() for o, g in zip(out, gs))
    loss.backward()
    return {
        "out": [o.detach() for o in out],
        "gout": list(gs),
        "gin": [x.grad for x in xs],
        "gparams": {k: p.grad for k, p in module.named_parameters() if p.grad is not None},
    }


def main():
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    gen = torch.Generator().manual_seed(1234)
    torch.manual_seed(0)
    fx = {}

    # --- known-answer test from the survey: 8-tuple x_i = (i+1)/4 through GeluD8 ---------------------------------
    kat_in = [torch.full((1, 1, 1), (i + 1) / 4.0) for i in range(8)]
    kat_out = d8_layers.GeluD8()(kat_in)
    fx["gelu_kat"] = {"in": kat_in, "out": list(kat_out)}

    # --- transforms and group actions ----------------------------------------------------------------------------
    xs8 = [torch.randn(2, 9, 4, generator=gen) for _ in range(8)]
    fx["transforms"] = {
        "in": xs8,
        "i2r": list(d8_utils.isotypic_to_regular_D8(xs8)),
        "r2i": list(d8_utils.regular_to_isotypic_D8(xs8)),
        "five": list(d8_utils.convert_8tuple_to_5tuple(xs8)),
    }
    img = torch.randn(2, 3, 6, 6, generator=gen)
    fx["actions"] = {
        "img": img, "xs8": xs8,
        "image": {g: d8_utils.image_space_group_action(g, img) for g in d8_utils.group_elements},
        "isotypic": {g: list(d8_utils.isotypic_group_action(g, xs8)) for g in d8_utils.group_elements},
        "spatial_isotypic": {g: list(d8_utils.spatial_and_isotypic_group_action(g, tuple(xs8)))
                             for g in d8_utils.group_elements},
    }

    # --- layers ---------------------------------------------------------------------------------------------------
    B, N = 2, 10
    lin = d8_layers.LinearD8(64, 128)
    randomize(lin, gen)
    fx["linear_d8"] = {"sd": lin.state_dict(), "in": list(rand5(B, N, 8, gen))}
    fx["linear_d8"].update(with_grads(lin, fx["linear_d8"]["in"], gen))

    ln = d8_layers.LayerNormD8(64)
    randomize(ln, gen)
    xs = tuple(x * 1.7 + 0.4 for x in rand5(B, N, 8, gen))
    fx["layernorm_d8"] = {"sd": ln.state_dict(), "in": list(xs)}
    fx["layernorm_d8"].update(with_grads(ln, xs, gen))

    gelu = CpuGeluD8()
    xs = rand5(B, N, 8, gen)
    fx["gelu_d8"] = {"in": list(xs)}
    fx["gelu_d8"].update(with_grads(gelu, xs, gen))

    mlp = swap_gelu(d8_layers.MlpD8(64, 256))
    randomize(mlp, gen, std=0.2)
    xs = rand5(B, N, 8, gen)
    fx["mlp_d8"] = {"sd": mlp.state_dict(), "in": list(xs)}
    fx["mlp_d8"].update(with_grads(mlp, xs, gen))

    attn = d8_layers.AttentionD8(128, num_heads=2, qkv_bias=True)
    randomize(attn, gen, std=0.15)
    xs = rand5(B, 17, 16, gen)
    fx["attention_d8"] = {"sd": attn.state_dict(), "in": list(xs), "num_heads": 2}
    fx["attention_d8"].update(with_grads(attn, xs, gen))

    blk = swap_gelu(d8_layers.Layer_scale_init_BlockD8(128, 2, qkv_bias=True))
    randomize(blk, gen, std=0.12)
    xs = rand5(B, 17, 16, gen)
    fx["block_deit_d8"] = {"sd": blk.state_dict(), "in": list(xs), "num_heads": 2}
    fx["block_deit_d8"].update(with_grads(blk, xs, gen))

    blk2 = swap_gelu(d8_layers.BlockD8(128, 2, qkv_bias=True, init_values=0.5))
    randomize(blk2, gen, std=0.12)
    xs = rand5(B, 17, 16, gen)
    fx["block_dinov2_d8"] = {"sd": blk2.state_dict(), "in": list(xs), "num_heads": 2}
    fx["block_dinov2_d8"].update(with_grads(blk2, xs, gen))

    inv = PowerSpectrumInvariant(64)
    xs = tuple(x.clone().requires_grad_(True) for x in rand5(B, N, 8, gen))
    y = inv(xs)
    gy = torch.randn(y.shape, generator=gen)
    (y * gy).sum().backward()
    fx["power_spectrum"] = {"in": [x.detach() for x in xs], "out": y.detach(), "gout": gy, "gin": [x.grad for x in xs]}

    # --- dense block (deit/vit.py) ----------------------------------------------------------------------------------
    dblk = Layer_scale_init_Block(64, 2, qkv_bias=True)
    randomize(dblk, gen, std=0.15)
    x = torch.randn(2, 9, 64, generator=gen, requires_grad=True)
    y = dblk(x)
    gy = torch.randn(y.shape, generator=gen)
    (y * gy).sum().backward()
    fx["dense_block_deit"] = {"sd": dblk.state_dict(), "in": x.detach(), "out": y.detach(), "gout": gy,
                              "gin": x.grad, "gparams": {k: p.grad for k, p in dblk.named_parameters()},
                              "num_heads": 2}

    # --- whole models (tiny) ------------------------------------------------------------------------------------------
    for tag, kwargs in {
        "hybrid": dict(invariant=False),
        "invariant": dict(invariant=True),
    }.items():
        model = OcticVisionTransformer(img_size=64, patch_size=16, embed_dim=64, depth=4, num_heads=2, num_classes=10,
                                       qkv_bias=True, standard_block_layers=Layer_scale_init_Block,
                                       octic_block_layers=d8_layers.Layer_scale_init_BlockD8, **kwargs)
        swap_gelu(model)
        randomize(model, gen, std=0.1)
        model.eval()
        img = torch.randn(2, 3, 64, 64, generator=gen)
        with torch.no_grad():
            xs_embed = model.patch_embed(img)
            logits = model(img)
            # octic trunk output (model.py:172-194), for the equivariance/parity tests of the trunk
            pos = d8_utils.convert_8tuple_to_5tuple(d8_utils.isotypic_dim_interpolation(model.pos_embed, dim=0))
            xs = tuple(x + v.flatten(0, 1) for x, v in zip(xs_embed, pos))
            cls = tuple(model.cls_token[i].expand(2, *model.cls_token[i].shape[1:]) for i in range(5))
            xs = tuple(torch.cat((cls[i], xs[i]), dim=1) for i in range(5))
            tokens0 = [t.clone() for t in xs]
            for b in model.blocks[:model.octic_equi_break_layer]:
                xs = b(xs)
        # gradients of a scalar loss w.r.t. a few parameters (training parity)
        model.train()
        logits_t = model(img)
        tgt = torch.randn(logits_t.shape, generator=gen)
        (logits_t * tgt).sum().backward()
        gsel = {k: p.grad.clone() for k, p in model.named_parameters()
                if p.grad is not None and (k.startswith("blocks.0.") or k.startswith("blocks.3.") or
                                           k.startswith("patch_embed") or k.startswith("pos_embed") or
                                           k.startswith("cls_token") or k.startswith("head") or k.startswith("invariant_proj"))}
        fx[f"model_{tag}"] = {
            "sd": {k: v.clone() for k, v in model.state_dict().items()},
            "cfg": dict(img_size=64, patch=16, embed_dim=64, depth=4, num_heads=2, num_classes=10, **kwargs),
            "img": img, "patch_embed": list(xs_embed), "tokens0": tokens0, "trunk": list(xs), "logits": logits,
            "loss_weight": tgt, "gparams": gsel,
        }

    # timm-default blocks (BlockD8 + timm Block): config 1 of BASELINE.json builds the model this way
    model = OcticVisionTransformer(img_size=32, patch_size=8, embed_dim=64, depth=2, num_heads=2, num_classes=5,
                                   init_scale=0.7)
    swap_gelu(model)
    randomize(model, gen, std=0.1)
    model.eval()
    img = torch.randn(2, 3, 32, 32, generator=gen)
    with torch.no_grad():
        logits = model(img)
    fx["model_timm_default"] = {"sd": {k: v.clone() for k, v in model.state_dict().items()},
                                "cfg": dict(img_size=32, patch=8, embed_dim=64, depth=2, num_heads=2, num_classes=5),
                                "img": img, "logits": logits}

    total = 0
    for name, obj in fx.items():
        path = out_dir / f"{name}.pt"
        torch.save(obj, path)
        total += path.stat().st_size
        print(f"{name:22s} {path.stat().st_size / 1024:8.1f} KiB")
    print(f"total {total / 1e6:.2f} MB in {out_dir}")


if __name__ == "__main__":
    main()
