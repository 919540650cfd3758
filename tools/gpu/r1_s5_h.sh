#!/usr/bin/env bash
# gamma-folded layer-scale backward: new kernel tests, all tests, events, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "layernorm or gamma_folded or layerscale" 2>&1 | tail -25 )
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
cat gpurun_out/events_b128.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
