#!/usr/bin/env bash
# zero pool + NCCL AVG: tests, bench (1 GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 | cut -c1-200 )
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; grep -v Warn gpurun_out/bench.err | tail -2 | cut -c1-200
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b64.csv python tools/profile_step.py --batch 64 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_b64.csv 12
