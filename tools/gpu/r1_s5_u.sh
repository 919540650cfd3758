#!/usr/bin/env bash
# packed-pair GELU epilogues: microbench, tests, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/microbench_ops.py --batch 128 --only dense_fc1_gelu,dense_fc2_dgrad,gelu_d8,dense_qkv 2>&1 | tail -6
( timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 | cut -c1-200 )
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-330 gpurun_out/bench.json
timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
cat gpurun_out/events_b128.txt
