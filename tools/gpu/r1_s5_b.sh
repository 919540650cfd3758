#!/usr/bin/env bash
# quick check: full GPU tests, bench, event table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
cat gpurun_out/events_b128.txt
