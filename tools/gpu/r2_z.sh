#!/usr/bin/env bash
# round 2, call Z: sparse-map front end -- model tests (goldens: patch embed, tokens, parameter gradients), launch list, bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_dinov2.py tests/test_gpu_parity_budget.py -x -q -m gpu > gpurun_out/r2z_tests_model.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2z_tests_model.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2z_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d["gpu_launches"])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_b64.csv python tools/profile_step.py --batch 64 > /dev/null 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2z_launches_b64.csv 60 > gpurun_out/r2z_launches_b64.txt 2>&1; grep -v "octic::" gpurun_out/r2z_launches_b64.txt | head -30
