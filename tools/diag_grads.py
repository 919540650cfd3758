"""Diagnostic: per-parameter gradient error of the GPU path vs the reference goldens (tiny whole models)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from octic_vits_b200 import layers as L
from octic_vits_b200.model import OcticVisionTransformer

for tag in ("hybrid", "invariant"):
    fx = torch.load(ROOT / "tests" / "golden" / f"model_{tag}.pt", weights_only=False)
    cfg = fx["cfg"]
    model = OcticVisionTransformer(img_size=64, patch_size=16, embed_dim=64, depth=4, num_heads=2, num_classes=10,
                                   qkv_bias=True, invariant=cfg["invariant"], standard_block_layers=L.Layer_scale_init_Block,
                                   octic_block_layers=L.Layer_scale_init_BlockD8).cuda().train()
    model.load_state_dict(fx["sd"])
    out = model(fx["img"].cuda())
    (out * fx["loss_weight"].cuda()).sum().backward()
    P = dict(model.named_parameters())
    errs = {k: float((P[k].grad.cpu() - g).norm() / g.norm()) for k, g in fx["gparams"].items()}
    print(tag, "logits rel", float((out.detach().cpu() - fx["logits"]).norm() / fx["logits"].norm()))
    for k, v in sorted(errs.items(), key=lambda kv: -kv[1])[:12]:
        print(f"    {k:50s} {v:.3e}")
