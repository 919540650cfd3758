"""BASELINE.json configs[4] datapoint: DINOv2-style octic ViT-L/14 step -- teacher no-grad forward + student
forward+backward on 2B global crops (224 px, iBOT masks with ratio 0.1-0.5 on half of the samples), swept over
`octic_equi_break_layer` (number of leading octic blocks).  The DINO/iBOT heads and losses are out of scope (SURVEY
section 8): the student loss here is an MSE to the teacher's cls/patch tokens, which exercises the same backbone
gradients.  Device-timed with CUDA events; images/s counts the 2B global crops of one step.

    python tools/bench_dinov2.py [--batch 64] [--steps 5] [--layers 0,6,12,18,23]
    python tools/bench_dinov2.py --multicrop [--batch 16] [--local 8]      # the recipe's step shape: 2 global + 8 local crops
                                                                          # per image; crop by crop vs concatenated vs graphed
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from octic_vits_b200 import ops  # noqa: E402
from octic_vits_b200.dinov2_models import OcticDinoVisionTransformer  # noqa: E402
from octic_vits_b200.optim import FusedOptimizer  # noqa: E402
from octic_vits_b200.parallel import FlatGrads, GraphedStep  # noqa: E402


def flops_per_image(k, depth=24, D=1024, N=257, p=14):
    lin_oct, lin_std, attn = 12 * D * D * 3 / 16, 12 * D * D, 2 * N * D
    mac = k * N * (lin_oct + attn) + (depth - k) * N * (lin_std + attn) + (N - 1) * 3 * p * p * D
    return 2 * mac


def multicrop(args):
    """Student on [2B global 224 px crops (iBOT masks), local*B local 112 px crops], teacher on the global crops: the
    DINOv2 recipe's step shape (dinov2/train/ssl_meta_arch.py:122-330, heads / losses replaced by MSE stand-ins).
    Three ways to run the same step: crop by crop, crops concatenated (one launch of every per-token kernel, attention per
    crop resolution), and the concatenated step captured in one CUDA graph."""
    dev = torch.device("cuda", 0)
    B2, BL = 2 * args.batch, args.local * args.batch
    torch.manual_seed(0)
    k = int(args.layers.split(",")[0]) if "," not in args.layers else 12
    kw = dict(img_size=224, patch_size=14, embed_dim=1024, depth=24, num_heads=16, num_register_tokens=0,
              octic_equi_break_layer=k, dynamic_img_size=True)
    student = OcticDinoVisionTransformer(drop_path_rate=0.3, **kw).to(dev).train()
    teacher = OcticDinoVisionTransformer(**kw).to(dev).eval()
    teacher.load_state_dict(student.state_dict())
    for p in teacher.parameters():
        p.requires_grad_(False)
    fg = FlatGrads(student.parameters())
    inputs = {"global": torch.randn(B2, 3, 224, 224, device=dev), "local": torch.randn(BL, 3, 112, 112, device=dev),
              "masks": torch.rand(B2, 256, device=dev) < 0.3}

    def forward_loss(inp):
        with torch.no_grad():
            t = teacher(inp["global"], is_training=True)
        sg, sl = student([inp["global"], inp["local"]], masks=[inp["masks"], None], is_training=True)
        m = inp["masks"].unsqueeze(-1).float()
        tc = t["x_norm_clstoken"].mean(0, keepdim=True)
        return ((((sg["x_norm_patchtokens"] - t["x_norm_patchtokens"]) ** 2) * m).sum() / m.sum().clamp_min(1.0) / 1024
                + ((sg["x_norm_clstoken"] - tc) ** 2).mean() + ((sl["x_norm_clstoken"] - tc) ** 2).mean())

    def eager():
        ops.begin_step()
        fg.zero()
        loss = forward_loss(inputs)
        loss.backward()
        return loss

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    rows = {}
    student.concat_crops = False
    rows["eager_crop_by_crop_ms"] = timed(eager)
    student.concat_crops = True
    rows["eager_concatenated_ms"] = timed(eager)
    step = GraphedStep(student, fg, inputs, forward_loss)
    rows["graph_captured"] = bool(step.graphed)
    if not step.graphed:
        rows["capture_error"] = getattr(step, "capture_error", "")
    rows["graphed_concatenated_ms"] = timed(lambda: step(**inputs))
    rows["crops_per_step"] = B2 + BL
    rows["crops_per_s_graphed"] = (B2 + BL) / rows["graphed_concatenated_ms"] * 1e3
    print(json.dumps({"config": f"DINOv2 octic ViT-L/14 (break layer {k}), teacher fwd on {B2} global crops + student fwd+bwd "
                                f"on {B2} global (224 px, iBOT masks) + {BL} local (112 px) crops, drop_path 0.3", **rows}))
    step.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--multicrop", action="store_true")
    ap.add_argument("--local", type=int, default=8, help="local crops per image (multicrop mode)")
    ap.add_argument("--batch", type=int, default=64, help="B: images per GPU; a step sees 2B global crops")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--layers", default="0,6,12,18,23")
    args = ap.parse_args()
    if args.multicrop:
        return multicrop(args)
    dev = torch.device("cuda", 0)
    B2 = 2 * args.batch
    torch.manual_seed(0)
    x = torch.randn(B2, 3, 224, 224, device=dev)
    masks = torch.zeros(B2, 256, dtype=torch.bool, device=dev)
    for i in range(0, B2, 2):                                   # half of the samples, ratio uniform in [0.1, 0.5]
        r = 0.1 + 0.4 * float(torch.rand(()))
        masks[i, torch.randperm(256, device=dev)[:int(256 * r)]] = True
    rows = []
    for k in [int(v) for v in args.layers.split(",")]:
        kw = dict(img_size=224, patch_size=14, embed_dim=1024, depth=24, num_heads=16, num_register_tokens=0,
                  octic_equi_break_layer=k)
        student = OcticDinoVisionTransformer(drop_path_rate=0.3, **kw).to(dev).train()
        teacher = OcticDinoVisionTransformer(**kw).to(dev).eval()
        teacher.load_state_dict(student.state_dict())
        for p in teacher.parameters():
            p.requires_grad_(False)
        fg = FlatGrads(student.parameters())
        opt = FusedOptimizer(student, fg, kind="adamw", lr=1e-4, weight_decay=0.04, ema=(teacher, 0.994))

        def step():
            ops.begin_step()
            fg.zero()
            with torch.no_grad():
                t = teacher(x, is_training=True)
            s = student(x, masks=masks, is_training=True)
            loss = (s["x_norm_clstoken"] - t["x_norm_clstoken"]).pow(2).mean() + \
                (s["x_norm_patchtokens"] - t["x_norm_patchtokens"]).pow(2).mean()
            loss.backward()
            opt.step()                                           # AdamW + teacher EMA, one launch
            return loss
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        fl = flops_per_image(k) * B2 * 4                         # teacher 1x + student 3x forward FLOPs
        rows.append({"octic_equi_break_layer": k, "ms_per_step": ms, "global_crops_per_s": B2 / ms * 1e3,
                     "model_tflops": fl / ms / 1e9, "loss": float(loss), "gflop_fwd_per_image": flops_per_image(k) / 1e9})
        print(json.dumps(rows[-1]), flush=True)
        del student, teacher, fg, opt
        torch.cuda.empty_cache()
    print(json.dumps({"config": "DINOv2 octic ViT-L/14, teacher fwd + student fwd+bwd + fused AdamW/EMA, eager (no graph), "
                                f"2B = {B2} global crops, drop_path 0.3, iBOT masks on half of the samples", "rows": rows}))


if __name__ == "__main__":
    main()
