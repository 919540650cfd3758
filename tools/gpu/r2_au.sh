#!/usr/bin/env bash
# round 2, call AU: memory of the graphed step at 256 images per GPU (who holds 85 GB outside the graph pool?)
set -u
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | grep -v Warning | tail -20
import gc, torch, sys
sys.path.insert(0, ".")
from octic_vits_b200.deit_models import create_model
from octic_vits_b200.parallel import FlatGrads, GraphedTrainStep
dev = torch.device("cuda", 0)
B = 256
model = create_model("hybrid_deit_huge_patch14", num_classes=1000).to(dev).train()
fg = FlatGrads(model.parameters())
gb = lambda: (torch.cuda.memory_allocated() / 2**30, torch.cuda.memory_reserved() / 2**30)
print("after model", gb())
img = torch.randn(B, 3, 224, 224, device=dev); tgt = torch.randint(0, 1000, (B,), device=dev)
loss = torch.nn.functional.cross_entropy(model(img), tgt); loss.backward()
torch.cuda.synchronize(); print("after one eager step (loss alive)", gb())
del loss; print("after del loss", gb())
gc.collect(); print("after gc", gb())
torch.cuda.empty_cache(); print("after empty_cache", gb())
try:
    step = GraphedTrainStep(model, fg, img.shape)
    print("graphed", step.graphed, getattr(step, "capture_error", "")[:200], gb())
    l = step(img, tgt); torch.cuda.synchronize(); print("replay ok", float(l), gb())
except Exception as e:
    print("FAILED", repr(e)[:300], gb())
PY
