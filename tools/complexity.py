"""The reference's throughput protocol (experiments/complexity.py:13-103: inference, batch 64, 10 warm-up + 100 timed
iterations, params / throughput / FLOPs / peak memory table) for the octic models of this package.

    python tools/complexity.py [--models a,b] [--batch 64] [--graph] [--dry-run]

Differences from the reference script, on purpose: timing uses CUDA events around each iteration instead of
time.time() + synchronize (mean AND median are printed); FLOPs are the analytic count of SURVEY.md section 8d
(FLOP = 2 MAC; fvcore is not installed); bf16 is the compute type of the kernels, so there is no --amp switch;
--graph replays the forward from a CUDA graph instead of --compile.  --dry-run builds the models and prints the static
columns only (no GPU needed)."""
import argparse
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from octic_vits_b200.deit_models import create_model  # noqa: E402

WARMUP_ITERATIONS = 10
NUM_ITERATIONS = 100
MODELS = ["d8_inv_early_deit_huge_patch14", "hybrid_deit_huge_patch14", "d8_inv_early_deit_large_patch16",
          "hybrid_deit_large_patch16"]
HEADERS = ["Model Name", "Params (10^6)", "Throughput (im/s)", "median ms", "GFLOP / image", "model TFLOP/s", "Peak Mem (MB)"]


def gflop_per_image(model) -> float:
    D, depth, k = model.embed_dim, len(model.blocks), model.octic_equi_break_layer
    p = model.patch_embed.patch_size[0]
    N = model.patch_embed.num_patches + (0 if model.global_pool else 1)
    lin_oct, lin_std, attn = 12 * D * D * 3 / 16, 12 * D * D, 2 * N * D
    mac = k * N * (lin_oct + attn) + (depth - k) * N * (lin_std + attn) + (N - 1) * 3 * p * p * D + D * model.num_classes
    if model.invariant:
        mac += N * (6 * D // 8) * D
    return 2 * mac / 1e9


@torch.no_grad()
def measure(name, batch, use_graph, dry):
    model = create_model(name).eval()
    params = sum(p.numel() for p in model.parameters())
    gf = gflop_per_image(model)
    if dry:
        return [name, f"{params / 1e6:.4f}", "-", "-", f"{gf:.1f}", "-", "-"]
    dev = torch.device("cuda", 0)
    model.to(dev)
    img = torch.randn(batch, 3, 224, 224, device=dev)
    torch.cuda.reset_peak_memory_stats()
    model(img)
    torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated(dev) / (1024 * 1024)
    run = lambda: model(img)
    if use_graph:
        for _ in range(2):
            model(img)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            model(img)
        run = g.replay
    for _ in range(WARMUP_ITERATIONS):
        run()
    torch.cuda.synchronize()
    times = []
    for _ in range(NUM_ITERATIONS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    mean, med = statistics.mean(times), statistics.median(times)
    del model
    torch.cuda.empty_cache()
    return [name, f"{params / 1e6:.4f}", f"{batch / mean * 1e3:.0f}", f"{med:.2f}", f"{gf:.1f}",
            f"{batch * gf / mean:.0f}", f"{peak:.0f}"]


def main():
    ap = argparse.ArgumentParser(description="Compute model complexity (reference protocol)")
    ap.add_argument("--models", default=",".join(MODELS))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--graph", action="store_true", help="replay the forward from a CUDA graph")
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args()
    rows = [measure(n, args.batch, args.graph, args.dry_run) for n in args.models.split(",")]
    try:
        from tabulate import tabulate
        print(tabulate(rows, headers=HEADERS, tablefmt="grid"))
    except ImportError:
        print(HEADERS)
        for r in rows:
            print(r)


if __name__ == "__main__":
    main()
