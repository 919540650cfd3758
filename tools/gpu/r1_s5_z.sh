#!/usr/bin/env bash
# BASELINE.json configs[1]: hybrid ViT-H/14 bf16 inference, batch 256, one B200 (CUDA-event table of one forward)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --batch 256 --infer --events > gpurun_out/events_infer_b256.txt 2>&1
cat gpurun_out/events_infer_b256.txt
timeout 300 python - <<'PY'
import torch, sys, time
sys.path.insert(0, '.')
from octic_vits_b200.deit_models import create_model
m = create_model("hybrid_deit_huge_patch14", num_classes=1000).cuda().eval()
x = torch.randn(256, 3, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(3): m(x)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"inference b256 (CUDA graph): {ms:.2f} ms/forward = {256/ms*1e3:.0f} img/s = {256/ms*1e3*203.2e9/1e12:.0f} model-TFLOP/s")
PY
