"""Times optim.FusedOptimizer.step() on the headline model's parameters (hybrid ViT-H/14, 356 M fp32 parameters) with
CUDA events and reports achieved HBM GB/s against MEASURED_PEAKS.json.  Algorithmic bytes per parameter element
(include/octic_b200.h): LAMB = 4 (grad norm) + 28 (stage 1) + 12 (stage 2) = 44 B, AdamW = 28 B (+4 with clipping);
EMA adds 8 B."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from octic_vits_b200.deit_models import create_model  # noqa: E402
from octic_vits_b200.optim import FusedOptimizer  # noqa: E402
from octic_vits_b200.parallel import FlatGrads  # noqa: E402

peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
dev = torch.device("cuda", 0)
model = create_model("hybrid_deit_huge_patch14", num_classes=1000).to(dev)
ema_model = create_model("hybrid_deit_huge_patch14", num_classes=1000).to(dev)
fg = FlatGrads(model.parameters())
n = sum(p.numel() for p in fg.params)
print(f"parameters: {n / 1e6:.1f} M in {len(fg.params)} tensors, flat buffer {fg.flat.numel() * 4 / 1e9:.2f} GB")
CASES = (("lamb", False, 44), ("lamb", True, 52), ("adamw", False, 28), ("adamw", True, 36))
if len(sys.argv) > 1 and sys.argv[1] == "--ncu":        # one LAMB+EMA and one AdamW+EMA step under ncu
    for kind in ("lamb", "adamw"):
        opt = FusedOptimizer(model, fg, kind=kind, lr=1e-4, weight_decay=0.05, ema=(ema_model, 0.999))
        fg.flat.normal_(std=1e-3)
        opt.step()
        torch.cuda.synchronize()
    sys.exit(0)
for kind, ema, bpe in CASES:
    opt = FusedOptimizer(model, fg, kind=kind, lr=1e-4, weight_decay=0.05, ema=(ema_model, 0.999) if ema else None)
    fg.flat.normal_(std=1e-3)
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = n * bpe / (ms * 1e-3) / 1e9
    print(f"{kind:5s} ema={ema!s:5s}: {ms:6.3f} ms/step, {opt.nchunks} chunks, algorithmic {n * bpe / 1e9:.2f} GB -> "
          f"{gbs:6.0f} GB/s = {gbs / peaks['hbm_gbs']:.2f} of measured HBM peak ({peaks['hbm_gbs']:.0f} GB/s)")
    del opt
