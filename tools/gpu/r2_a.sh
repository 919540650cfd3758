#!/usr/bin/env bash
# round 2, call A: the unmodified reference on this B200 + compute-sanitizer over the kernel tests.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python tools/bench_reference_gpu.py --out gpurun_out/reference_gpu.json > gpurun_out/r2a_reference.log 2>&1
echo "reference rc=$?"
tail -5 gpurun_out/r2a_reference.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm or attention or linear_d8" > gpurun_out/r2a_memcheck.log 2>&1
echo "memcheck rc=$?"
tail -8 gpurun_out/r2a_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "test_attention and 17-2-64 or test_attention and 129-2-80 or test_gemm_dense_bf16 and 257-64-48 or test_linear_d8_forward and 2-50-64-128 or test_gemm_wgrad and 512-128-64" > gpurun_out/r2a_racecheck.log 2>&1
echo "racecheck rc=$?"
tail -8 gpurun_out/r2a_racecheck.log
