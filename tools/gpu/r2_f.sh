#!/usr/bin/env bash
# round 2, call F: attention bwd (tail stores moved out of the piece loop; timing variants) + weight-stationary octic GEMM.
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r2f_tests_attention.log 2>&1; echo "attention tests rc=$?"
tail -2 gpurun_out/r2f_tests_attention.log
for v in attn_trace attn_trace_nots; do
  timeout 60 build/$v 20 b 1 > gpurun_out/r2f_${v}_b20.txt 2>&1; echo "$v rc=$?"; head -1 gpurun_out/r2f_${v}_b20.txt
  timeout 60 build/$v 128 b 1 > gpurun_out/r2f_${v}_b128.txt 2>&1; head -1 gpurun_out/r2f_${v}_b128.txt
done
grep -A8 "slot 1" gpurun_out/r2f_attn_trace_b20.txt | head -9
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm or linear_d8 or gamma" > gpurun_out/r2f_tests_gemm.log 2>&1; echo "gemm tests rc=$?"
tail -4 gpurun_out/r2f_tests_gemm.log
timeout 200 python tools/microbench_ops.py --batch 128 --only d8_,attn > gpurun_out/r2f_microbench_ws.txt 2>&1; echo "microbench ws rc=$?"
cat gpurun_out/r2f_microbench_ws.txt | tail -25
OCTIC_GEMM_WS=0 timeout 200 python tools/microbench_ops.py --batch 128 --only d8_ > gpurun_out/r2f_microbench_nows.txt 2>&1
cat gpurun_out/r2f_microbench_nows.txt | tail -22
