#!/usr/bin/env bash
# round 2, call AH: graphed DINOv2 step (parallel.GraphedStep) + multi-crop step bench (crop by crop / concatenated / graphed)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dinov2.py -x -q -m gpu > gpurun_out/r2ah_tests_dinov2.log 2>&1; echo "dinov2 tests rc=$?"; tail -12 gpurun_out/r2ah_tests_dinov2.log
timeout 400 python tools/bench_dinov2.py --multicrop --batch 16 --local 8 --steps 5 > gpurun_out/r2ah_bench_dinov2_multicrop.txt 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/r2ah_bench_dinov2_multicrop.txt
