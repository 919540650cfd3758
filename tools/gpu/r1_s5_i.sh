#!/usr/bin/env bash
# CUDA-graphed step: tests, bench with and without the graph
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err
cat gpurun_out/bench_nograph.json; tail -3 gpurun_out/bench_nograph.err
timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
cat gpurun_out/events_b128.txt
