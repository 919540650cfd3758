#!/usr/bin/env bash
# fused gradient accumulation: all tests, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -12 ) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-330 gpurun_out/bench.json; grep -v Warning gpurun_out/bench.err | tail -3 | cut -c1-300
