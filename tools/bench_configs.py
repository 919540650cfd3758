"""Datapoints for the BASELINE.json configs that are not the headline line of bench.py (one GPU, synthetic data,
random-init weights, device-timed with CUDA events, steps replayed from a CUDA graph):

  configs[1]  hybrid octic ViT-H/14 bf16 inference, batch 256
  configs[2]  invariant octic ViT-L/16 (d8_inv_early_deit_large_patch16) DeiT-III fwd+bwd, batch 256 per GPU
  configs[3]  hybrid ViT-H/14 DeiT-III training step incl. the fused LAMB update (drop_path 0.5 as in the recipe)

    python tools/bench_configs.py [--steps 6] [--only 1,2,3]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from octic_vits_b200.deit_models import create_model  # noqa: E402
from octic_vits_b200.optim import FusedOptimizer  # noqa: E402
from octic_vits_b200.parallel import FlatGrads, GraphedTrainStep  # noqa: E402

DEV = torch.device("cuda", 0)


def flops_fwd(D, depth, N, p, invariant, ncls=1000):
    k = depth // 2
    lin_oct, lin_std, attn = 12 * D * D * 3 / 16, 12 * D * D, 2 * N * D
    mac = k * N * (lin_oct + attn) + (depth - k) * N * (lin_std + attn) + (N - 1) * 3 * p * p * D + D * ncls
    if invariant:
        mac += N * (6 * D // 8) * D
    return 2 * mac


def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def inference(steps):
    B = 256
    m = create_model("hybrid_deit_huge_patch14", num_classes=1000).to(DEV).eval()
    x = torch.randn(B, 3, 224, 224, device=DEV)
    with torch.no_grad():
        for _ in range(3):
            m(x)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            m(x)
        ms = timed(g.replay, steps)
    fl = flops_fwd(1280, 32, 257, 14, False)
    return {"config": "configs[1] hybrid ViT-H/14 bf16 inference, batch 256, CUDA graph", "ms": ms,
            "images_per_s": B / ms * 1e3, "model_tflops": B * fl / ms / 1e9}


def train(name, B, D, depth, N, p, invariant, steps, drop_path=0.0, optimizer=None, tag=""):
    m = create_model(name, num_classes=1000, drop_path_rate=drop_path).to(DEV).train()
    fg = FlatGrads(m.parameters())
    opt = FusedOptimizer(m, fg, kind=optimizer, lr=1e-3, weight_decay=0.05) if optimizer else None
    step = GraphedTrainStep(m, fg, (B, 3, 224, 224), optimizer=opt)
    x = torch.randn(B, 3, 224, 224, device=DEV)
    t = torch.randint(0, 1000, (B,), device=DEV)
    for _ in range(2):
        step(x, t)
    ms = timed(lambda: step(x, t), steps)
    fl = 3 * flops_fwd(D, depth, N, p, invariant)
    return {"config": tag, "graphed": step.graphed, "ms": ms, "images_per_s": B / ms * 1e3,
            "model_tflops": B * fl / ms / 1e9, "optimizer": optimizer, "drop_path": drop_path}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--only", default="1,2,3")
    args = ap.parse_args()
    todo = {int(v) for v in args.only.split(",")}
    if 1 in todo:
        print(json.dumps(inference(args.steps)), flush=True)
        torch.cuda.empty_cache()
    if 2 in todo:
        print(json.dumps(train("d8_inv_early_deit_large_patch16", 256, 1024, 24, 197, 16, True, args.steps,
                               tag="configs[2] invariant octic ViT-L/16 DeiT-III fwd+bwd, batch 256")), flush=True)
        torch.cuda.empty_cache()
    if 3 in todo:
        print(json.dumps(train("hybrid_deit_huge_patch14", 128, 1280, 32, 257, 14, False, args.steps, drop_path=0.5,
                               optimizer="lamb", tag="configs[3] hybrid ViT-H/14 DeiT-III step: fwd+bwd (drop_path 0.5) + "
                                                     "fused LAMB, batch 128")), flush=True)


if __name__ == "__main__":
    main()
