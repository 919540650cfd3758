#!/usr/bin/env bash
# round 2, call S: bf16-packed epilogue staging -- correctness + A/B timing
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r2s_tests_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -3 gpurun_out/r2s_tests_kernels.log
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "golden or config1 or drop_path" > gpurun_out/r2s_tests_model.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2s_tests_model.log
OPS=dense_qkv,dense_fc1_gelu,dense_fc2_resid,dense_proj_resid,dense_proj_plain,dense_fc1_dgrad,d8_qkv,d8_fc1,d8_fc2_resid,d8_proj_resid,d8_fc1_dgrad,d8_fc2_dgrad
echo "== packed (default)"; timeout 200 python tools/microbench_ops.py --batch 128 --only $OPS 2>&1 | tail -13 | tee gpurun_out/r2s_microbench_packed.txt
echo "== OCTIC_GEMM_PACKED=0"; OCTIC_GEMM_PACKED=0 timeout 200 python tools/microbench_ops.py --batch 128 --only $OPS 2>&1 | tail -13 | tee gpurun_out/r2s_microbench_fp32staged.txt
