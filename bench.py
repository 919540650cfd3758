#!/usr/bin/env python
"""bench.py -- headline benchmark of the octic ViT hot path (contract: see the task brief / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
                    [--optimizer none|lamb|adamw] [--overlap] [--drop-path P] [--no-graph] [--no-cpu-baseline]

Metric (BASELINE.json): hybrid octic ViT-H/14 (DeiT-III: embed 1280, depth 32, heads 16, patch 14, 224 px) images/s,
forward + backward (+ NCCL gradient all-reduce when N > 1), bf16 compute / fp32 residual, synthetic images, random-init
weights.  One process per GPU (torchrun for N > 1); the batch is sharded over ranks (weak scaling: B images per GPU).

The step is the public-API parallel.GraphedTrainStep: bf16 weight re-pack + forward + loss + backward replayed from one
CUDA graph.  `--optimizer` adds the fused parameter update (optim.FusedOptimizer) to every step; `--overlap` (N > 1)
moves the all-reduce of the dense half's gradients inside the graph, concurrent with the octic half's backward.

`--impl reference` times the reference algorithm on the host CPU cores (the oracle port of the reference PyTorch
path, oracle/octic_oracle.py -- the reference itself is pure PyTorch and is not present on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

MODEL = dict(name="hybrid_deit_huge_patch14", img=224, patch=14, dim=1280, depth=32, heads=16, classes=1000)


# ----------------------------------------------------------------------------------------------------------------
# algorithmic work: FLOP = 2 * MAC, backward = 2 x forward
# ----------------------------------------------------------------------------------------------------------------
def model_flops_per_image(fwd_bwd: bool = True) -> float:
    D, depth, p = MODEL["dim"], MODEL["depth"], MODEL["patch"]
    N = (MODEL["img"] // p) ** 2 + 1
    k = depth // 2
    lin_oct, lin_std, attn = 12 * D * D * 3 / 16, 12 * D * D, 2 * N * D
    mac = k * N * (lin_oct + attn) + (depth - k) * N * (lin_std + attn) + (N - 1) * 3 * p * p * D + D * MODEL["classes"]
    return 2 * mac * (3 if fwd_bwd else 1)


def load_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ----------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference PyTorch path on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(batch: int):
    from oracle import octic_oracle as O
    from octic_vits_b200.deit_models import create_model
    torch.manual_seed(0)
    model = create_model(MODEL["name"], num_classes=MODEL["classes"])          # parameters only (CPU tensors)
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "cls_token" not in k or k == "cls_token.0")
              for k, v in model.state_dict().items()}
    del model
    img = torch.randn(batch, 3, MODEL["img"], MODEL["img"])
    tgt = torch.randint(0, MODEL["classes"], (batch,))

    def step():
        for p in params.values():
            p.grad = None
        with torch.autocast("cpu", dtype=torch.bfloat16):
            logits = O.octic_vit_forward(img, params, patch=MODEL["patch"], depth=MODEL["depth"], num_heads=MODEL["heads"])
        loss = torch.nn.functional.cross_entropy(logits.float(), tgt)
        loss.backward()
        return loss.item()
    return step


def time_cpu_reference(batch: int, steps: int, warmup: int):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_reference_step_fn(batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 4
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    ips, dt, cores = time_cpu_reference(batch, steps, warmup)
    line = {
        "impl": "reference", "metric": "hybrid octic ViT-H/14 images/s fwd+bwd", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "hybrid_deit_huge_patch14 fwd+bwd, 224px, CPU oracle port of the reference PyTorch path "
                               "(bf16 autocast), bounded sample", "batch": batch},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"batch {batch}, {warmup} warm-up + {steps} timed fwd+bwd steps"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from octic_vits_b200 import _lib, functional as OF, ops
    from octic_vits_b200.deit_models import create_model

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if not _lib.load().octic_device_ok():
        raise SystemExit("bench.py: no sm_100 device -- the CUDA path is the only path")

    torch.manual_seed(1234 + rank)
    B = args.batch
    model = create_model(MODEL["name"], num_classes=MODEL["classes"], drop_path_rate=args.drop_path).to(dev).train()
    # one flat fp32 gradient buffer: .grad tensors are views, the all-reduce is a single NCCL call over NVLink
    from octic_vits_b200.parallel import FlatGrads, GraphedTrainStep
    fg = FlatGrads(model.parameters())
    flat = fg.flat
    opt = None
    if args.optimizer != "none":
        # fused parameter update (optim.FusedOptimizer: LAMB = the DeiT-III recipe, experiments/train_deit.py:42)
        from octic_vits_b200.optim import FusedOptimizer
        opt = FusedOptimizer(model, fg, kind=args.optimizer, lr=1e-3, weight_decay=0.05)

    img_dev = torch.randn(B, 3, MODEL["img"], MODEL["img"], device=dev)
    tgt_dev = torch.randint(0, MODEL["classes"], (B,), device=dev)
    img_host = torch.randn(B, 3, MODEL["img"], MODEL["img"]).pin_memory()
    tgt_host = torch.randint(0, MODEL["classes"], (B,)).pin_memory()

    def eager_step(img, tgt):
        ops.begin_step()
        flat.zero_()
        logits = model(img)
        loss = torch.nn.functional.cross_entropy(logits, tgt)
        loss.backward()
        fg.all_reduce()
        if opt is not None:
            opt.step()
        return loss

    # public-API step: fwd + loss + bwd captured in a CUDA graph (parallel.GraphedTrainStep), all-reduce after the replay
    overlap = False
    if world > 1 and args.overlap:
        # the dense half's gradients (88 % of the bytes) are all-reduced from an autograd hook at the bridge, on a side
        # stream inside the captured graph, while the octic half is still in backward
        from octic_vits_b200.parallel import install_early_allreduce
        overlap = install_early_allreduce(model, fg)
    gstep = GraphedTrainStep(model, fg, img_dev.shape, warmup=args.warmup, use_graph=not args.no_graph, optimizer=opt)
    overlap = overlap and model._bridge_grad_hook is not None
    step = gstep if gstep.graphed else eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    for _ in range(args.warmup):
        step(img_dev, tgt_dev)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(lambda: step(img_dev, tgt_dev), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # launches of our kernels inside the timed region: counted on one eager step (a graph replay launches the same nodes)
    from octic_vits_b200 import functional as OF
    if gstep.graphed:
        OF.bump_param_epoch()         # the graphed step re-packs the bf16 weights every replay: count those launches too
    _lib.STATS.reset()
    eager_step(img_dev, tgt_dev)
    torch.cuda.synchronize()
    launches = _lib.STATS.kernel_launches * args.steps

    # dominant kernel: the tcgen05 grouped GEMM.  Time every launch of it with CUDA events on the launching stream
    # during extra (untimed-for-the-headline) steps, so the headline number carries no event overhead.
    _lib.STATS.reset()
    _lib.STATS.profile_prefixes = ("octic_gemm_bf16", "octic_linear_d8_fwd", "octic_linear_d8_dgrad")
    eager_step(img_dev, tgt_dev)
    torch.cuda.synchronize()
    gemm_ms, gemm_flops, gemm_calls = _lib.STATS.collect()
    _lib.STATS.profile_prefixes = ()

    if gstep.graphed:
        gstep.stage(img_host, tgt_host)

    def e2e_step():
        if gstep.graphed:
            # public-API input pipeline: the step consumes the staged batch, the H2D copy of the NEXT batch (every step
            # copies its own inputs from pinned host memory) runs on the copy stream meanwhile, then the loss is read back
            loss = gstep.run()
            gstep.stage(img_host, tgt_host)
            return loss.item()
        img = img_host.to(dev, non_blocking=True)
        tgt = tgt_host.to(dev, non_blocking=True)
        return eager_step(img, tgt).item()
    e2e_step()
    ms_e2e = timed(e2e_step, max(2, args.steps // 2))

    if rank == 0:
        peaks = load_peaks()
        ips = world * B / (ms_dev * 1e-3)
        flops_img = model_flops_per_image(True)
        achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        cpu_ips, cpu_dt, cores = (None, None, None)
        if world == 1 and not args.no_cpu_baseline:
            cpu_ips, cpu_dt, cores = time_cpu_reference(2, 2, 1)
        line = {
            "metric": "hybrid octic ViT-H/14 images/s fwd+bwd", "value": ips, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "hybrid_deit_huge_patch14 (embed 1280, depth 32: 16 octic + 16 dense, heads 16, patch 14) "
                                   "DeiT-III training step: fwd + bwd" + (" + NCCL grad all-reduce" if world > 1 else ""),
                       "img": MODEL["img"], "batch_per_gpu": B, "global_batch": world * B, "tokens_per_image": 257,
                       "drop_path": args.drop_path, "parallelism": f"dp{world}",
                       "optimizer_step": False if opt is None else f"fused {args.optimizer} (3 launches/step, weight "
                                                                    "re-pack inside the graph)",
                       "cuda_graph": bool(gstep.graphed), "allreduce_overlap": bool(overlap),
                       "weight_repack_in_step": True,   # fp32 -> bf16 weight packs run every step (as autocast re-casts)
                       "l2": "activations per step (>50 GB) exceed the 126 MB L2; no explicit flush"},
            "model_tflops": ips * flops_img / 1e12,
            "tc_util_vs_sustained_peak": ips * flops_img / 1e12 / (world * peaks["tf_sustained"]),
            "roofline": {"kernel": "gemm_tn_kernel (tcgen05 grouped bf16 GEMM: LinearD8 fwd/dgrad + dense Linear fwd/dgrad)",
                         "bound": "tensor", "achieved": achieved_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": (achieved_tf / peaks["tf_sustained"]) if achieved_tf else None, "traffic": None,
                         "peak_source": f"{peaks['src']} (bf16 sustained)", "launches_timed": gemm_calls,
                         "share_of_step": gemm_ms / ms_dev,
                         # `achieved` averages 259 launches of many shapes, so there is no single per-launch traffic figure;
                         # ncu --set full per shape (profiles/r01_ncu_s5.md, batch 128): DRAM bytes vs algorithmic bytes
                         "traffic_samples": {
                             "dense qkv fwd (M=32896, N=3840, K=1280)": {"dram_bytes": 3.01e8, "algorithmic_bytes": 3.47e8},
                             "dense fc2 + residual (N=1280, K=5120)": {"dram_bytes": 8.79e8, "algorithmic_bytes": 7.71e8},
                             "octic fc2 + residual (LinearD8 5120 -> 1280)": {"dram_bytes": 7.35e8, "algorithmic_bytes": 7.62e8},
                             "octic qkv head-major (LinearD8 1280 -> 3840)": {"dram_bytes": 4.51e8, "algorithmic_bytes": 3.40e8}}},
            "cpu_baseline": {"value": cpu_ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": "batch 2, 1 warm-up + 2 timed fwd+bwd steps of the same model (oracle port, bf16 autocast)"},
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": img_host.numel() * 4 + tgt_host.numel() * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # a CUDA graph that captured NCCL kernels keeps the communicator alive: release it before tearing NCCL down
        # (with --overlap the process otherwise hangs in destroy_process_group, observed at 2 GPUs)
        if overlap:
            gstep.close()
            dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="images per GPU")
    ap.add_argument("--drop-path", type=float, default=0.0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the eager step instead of the captured CUDA graph")
    ap.add_argument("--overlap", action="store_true",
                    help="N > 1: all-reduce the dense half's gradients from inside the captured graph while the octic half "
                         "is still in backward (measured +1.0 %% at 2 GPUs; opt-in, see DESIGN.md section 5)")
    ap.add_argument("--optimizer", default="none", choices=["none", "lamb", "adamw"],
                    help="also run the fused parameter update every step (the headline metric is fwd+bwd)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
