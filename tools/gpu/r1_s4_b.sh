#!/usr/bin/env bash
# full GPU tests, pointwise microbench, bench, event table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 200 python tools/microbench_ops.py --batch 128 --only ln_,gelu,layerscale,colsum 2>&1 | tail -10 | tee gpurun_out/pointwise_bench.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
cat gpurun_out/events_b128.txt
