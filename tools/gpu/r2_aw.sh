#!/usr/bin/env bash
# round 2, call AW: by-products of the layer-norm backward ride on the gradient tensor (no module-global slot): hit rate + tests
set -u
mkdir -p gpurun_out
timeout 200 python tools/profile_step.py --batch 64 --events > gpurun_out/r2aw_events_b64.txt 2>&1; grep -E "step total|layerscale|layernorm" gpurun_out/r2aw_events_b64.txt
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_parity_budget.py tests/test_gpu_boundary.py -x -q -m gpu > gpurun_out/r2aw_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2aw_tests.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2aw_bench.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r2aw_bench.json'));print(d['value'], d['ms_per_step'], d['gpu_launches']/d['steps'])"
