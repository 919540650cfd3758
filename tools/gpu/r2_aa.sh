#!/usr/bin/env bash
# round 2, call AA: full GPU suite + sanitizer on the kernels changed this round
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2aa_tests_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2aa_tests_gpu.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -m gpu -k "gemm or linear_d8 or layernorm or golden or gamma" > gpurun_out/r2aa_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2aa_memcheck.log | head
python __graft_entry__.py smoke 2>&1 | tail -4
