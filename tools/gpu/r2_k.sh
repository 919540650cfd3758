#!/usr/bin/env bash
# round 2, call K: full GPU test suite + headline bench on the current tree.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_tests_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -4 gpurun_out/r2k_tests_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/r2k_bench.json
tail -3 gpurun_out/r2k_bench.err
