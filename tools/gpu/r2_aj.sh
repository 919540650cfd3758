#!/usr/bin/env bash
# round 2, call AJ: state of the round -- full GPU suite, smoke, default bench (both arms), launch list of one step, event table, inference
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2aj_tests_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2aj_tests_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2aj_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2aj_smoke.log | cut -c1-300
timeout 400 python bench.py > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2aj_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d["gpu_launches"], d["roofline"]["frac"], d.get("roofline_octic",{}).get("frac"), d.get("tc_util_vs_sustained_peak"))
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2aj_launches_b64.csv python tools/profile_step.py --batch 64 > /dev/null 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2aj_launches_b64.csv 60 > gpurun_out/r2aj_launches_b64.txt 2>&1; head -24 gpurun_out/r2aj_launches_b64.txt
timeout 200 python tools/profile_step.py --batch 128 --events > gpurun_out/r2aj_events_b128.txt 2>&1; echo "events rc=$?"; head -24 gpurun_out/r2aj_events_b128.txt
timeout 200 python tools/profile_step.py --batch 256 --infer --events > gpurun_out/r2aj_events_infer_b256.txt 2>&1; echo "infer events rc=$?"; head -12 gpurun_out/r2aj_events_infer_b256.txt
