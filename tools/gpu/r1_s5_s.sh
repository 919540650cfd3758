#!/usr/bin/env bash
# staged residual prefetch + aligned-split wgrad remainder: tests, microbench, bench, events
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | cut -c1-200 ) 2>&1 | tail -7
timeout 300 python tools/microbench_ops.py --batch 128 --only d8_proj_resid,d8_fc2_resid,dense_proj_resid,dense_fc2_resid,d8_fc1_wgrad,d8_qkv_wgrad,dense_proj_wgrad,dense_fc1_wgrad 2>&1 | tail -n +2
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-330 gpurun_out/bench.json
timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
head -9 gpurun_out/events_b128.txt
