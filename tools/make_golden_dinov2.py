"""Generate tests/golden/model_dinov2*.pt by running the REFERENCE OcticDinoVisionTransformer
(/root/reference/octic_vits/dinov2_models.py:40-260) on CPU.  Same stand-ins as tools/make_golden.py (timm shim, CPU
GeluD8 wrapper); kept as its own script so that the older fixtures keep their bits.

Build container only: `python tools/make_golden_dinov2.py`.

The reference cannot run list inputs without xformers (dinov2/layers/block.py:255-257); in eval
mode a list forward is, by construction of the block-diagonal mask, the per-crop forward of each element, so the
"list" golden is produced crop by crop.
"""
import sys
import warnings
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
warnings.filterwarnings("ignore")
from make_golden import ROOT, randomize, swap_gelu  # noqa: E402  (also puts the shim + reference on sys.path)

from octic_vits.dinov2_models import OcticDinoVisionTransformer  # noqa: E402


def build(gen, invariant, regs):
    torch.manual_seed(7)
    model = OcticDinoVisionTransformer(img_size=64, patch_size=16, embed_dim=64, depth=4, num_heads=2,
                                       num_register_tokens=regs, invariant=invariant)
    swap_gelu(model)
    randomize(model, gen, std=0.1)
    with torch.no_grad():                       # the frozen components stay zero, as in the reference
        for plist in (model.cls_token, model.mask_token, model.register_tokens or []):
            for i, p in enumerate(plist):
                if i > 0:
                    p.zero_()
    return model


def main():
    out_dir = ROOT / "tests" / "golden"
    gen = torch.Generator().manual_seed(4321)
    for tag, invariant, regs in (("model_dinov2", False, 2), ("model_dinov2_inv", True, 0)):
        model = build(gen, invariant, regs)
        model.eval()
        img = torch.randn(3, 3, 64, 64, generator=gen)
        img2 = torch.randn(2, 3, 64, 64, generator=gen)
        masks = torch.rand(3, 16, generator=gen) < 0.4
        masks[1] = False                        # one un-masked sample, as iBOT masks half of the batch
        with torch.no_grad():
            plain = model(img)                                        # forward(): cls token features
            feat = model(img, masks=masks, is_training=True)          # dict
            feat2 = model(img2, is_training=True)
            inter = model.get_intermediate_layers(img, n=1, reshape=False, return_class_token=True, norm=True)
            tokens0 = model.prepare_tokens_with_masks(img, masks)
        model.train()
        out = model(img, masks=masks, is_training=True)
        w_cls = torch.randn(out["x_norm_clstoken"].shape, generator=gen)
        w_patch = torch.randn(out["x_norm_patchtokens"].shape, generator=gen)
        ((out["x_norm_clstoken"] * w_cls).sum() + (out["x_norm_patchtokens"] * w_patch).sum()).backward()
        gsel = {k: p.grad.clone() for k, p in model.named_parameters()
                if p.grad is not None and (k.startswith(("blocks.0.", "blocks.3.", "patch_embed", "pos_embed", "cls_token",
                                                         "mask_token", "register_tokens", "norm", "invariant_proj")))}
        keep = ("x_norm_clstoken", "x_norm_regtokens", "x_norm_patchtokens", "x_prenorm")
        obj = {
            "sd": {k: v.clone() for k, v in model.state_dict().items()},
            "cfg": dict(img_size=64, patch_size=16, embed_dim=64, depth=4, num_heads=2, num_register_tokens=regs,
                        invariant=invariant),
            "img": img, "img2": img2, "masks": masks, "plain": plain,
            "feat": {k: feat[k] for k in keep}, "feat2": {k: feat2[k] for k in keep},
            "inter_patch": inter[0][0], "inter_cls": inter[0][1], "tokens0": list(tokens0),
            "w_cls": w_cls, "w_patch": w_patch, "gparams": gsel,
        }
        path = out_dir / f"{tag}.pt"
        torch.save(obj, path)
        print(f"{tag:22s} {path.stat().st_size / 1024:8.1f} KiB  keys={len(obj['sd'])}")


if __name__ == "__main__":
    main()
