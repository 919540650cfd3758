#!/usr/bin/env bash
# round 2, call C: attention backward with dedicated flush warps -- parity tests, timeline, microbench.
set -u
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention > gpurun_out/r2c_tests_attention.log 2>&1; echo "attention tests rc=$?"
tail -3 gpurun_out/r2c_tests_attention.log
timeout 60 build/attn_trace 20 b 1 > gpurun_out/r2c_trace_bwd_octic.txt 2>&1; echo "trace bwd rc=$?"
head -c 2500 gpurun_out/r2c_trace_bwd_octic.txt
timeout 120 python tools/microbench_ops.py --batch 128 --only attn > gpurun_out/r2c_microbench_attn.txt 2>&1; echo "microbench rc=$?"
tail -6 gpurun_out/r2c_microbench_attn.txt
timeout 120 python tools/microbench_ops.py --batch 128 --only attn --attn-layout 0 > gpurun_out/r2c_microbench_attn_dense.txt 2>&1
tail -4 gpurun_out/r2c_microbench_attn_dense.txt
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_optim.py -x -q -m gpu > gpurun_out/r2c_tests_model.log 2>&1; echo "model tests rc=$?"
tail -3 gpurun_out/r2c_tests_model.log
