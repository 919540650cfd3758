#!/usr/bin/env bash
# round 2, call AG: DINOv2 crop lists on concatenated rows (segmented attention) + the full GPU suite
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dinov2.py -x -q -m gpu > gpurun_out/r2ag_tests_dinov2.log 2>&1; echo "dinov2 tests rc=$?"; tail -15 gpurun_out/r2ag_tests_dinov2.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2ag_tests_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2ag_tests_gpu.log
