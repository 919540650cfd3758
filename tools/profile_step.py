"""One training (or inference) step of the headline model between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_step.py
Also usable stand-alone: prints a per-entry-point CUDA-event time table (each C-ABI call timed on its stream)."""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from octic_vits_b200 import _lib  # noqa: E402
from octic_vits_b200.deit_models import create_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="hybrid_deit_huge_patch14")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--infer", action="store_true")
ap.add_argument("--events", action="store_true", help="time every C-ABI call with CUDA events and print a table")
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = create_model(args.model, num_classes=1000).to(dev)
model.train(not args.infer)
img = torch.randn(args.batch, 3, 224, 224, device=dev)
tgt = torch.randint(0, 1000, (args.batch,), device=dev)


def step():
    if args.infer:
        with torch.no_grad():
            return model(img)
    for p in model.parameters():
        p.grad = None
    loss = torch.nn.functional.cross_entropy(model(img), tgt)
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()

if args.events:
    names = tuple(_lib.SIGNATURES)
    _lib.STATS.reset()
    _lib.STATS.profile_prefixes = names
    # keep the per-call name: wrap collect
    orig_call = _lib.call
    log = []

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    per = {}
    # events were appended in call order; rebuild names from the calls dict is lossy, so re-run with a name tap
    _lib.STATS.reset()
    tap = []
    def tapped(name, *a, **k):
        tap.append(name)
        return orig_call(name, *a, **k)
    import octic_vits_b200.ops as ops
    ops.call = tapped
    step()
    torch.cuda.synchronize()
    evs = _lib.STATS._events
    for name, (_n, a, b, f) in zip(tap, evs):
        ms = a.elapsed_time(b)
        d = per.setdefault(name, [0.0, 0, 0.0])
        d[0] += ms; d[1] += 1; d[2] += f
    print(f"step total {total:.2f} ms (batch {args.batch}, {'infer' if args.infer else 'train'})")
    tsum = sum(v[0] for v in per.values())
    for name, (ms, n, f) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        tf = f"{f / ms / 1e9:8.1f} TFLOP/s" if f else ""
        print(f"  {name:34s} {ms:9.3f} ms  {100 * ms / total:5.1f}%  calls {n:4d}  {tf}")
    print(f"  (sum of timed calls {tsum:.2f} ms; remainder = torch glue kernels + gaps)")
else:
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled one step")
