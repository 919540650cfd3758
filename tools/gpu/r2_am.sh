#!/usr/bin/env bash
# round 2, call AM: compute-sanitizer over the attention kernels changed this round (staged-dQ backward, one-pass forward)
set -u
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention and not recycle" > gpurun_out/r2am_memcheck_attention.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2am_memcheck_attention.log | head
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "test_attention and 257" > gpurun_out/r2am_racecheck_attention.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2am_racecheck_attention.log | head
