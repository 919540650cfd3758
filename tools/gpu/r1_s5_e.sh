#!/usr/bin/env bash
# epilogue policy experiment: register-direct vs staged epilogues, CTA pairs vs single CTAs, per op
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ONLY=d8_qkv,d8_proj,d8_fc1,d8_fc2,dense_qkv,dense_proj_resid,dense_fc1_gelu,dense_fc2_resid,dense_proj_plain,dense_fc2_dgrad,dense_fc1_dgrad
for cfg in "31 2" "0 2" "31 1" "0 1"; do
  set -- $cfg
  echo "## direct_mask=$1 ncta=$2"
  OCTIC_GEMM_DIRECT=$1 OCTIC_GEMM_NCTA=$2 timeout 300 python tools/microbench_ops.py --batch 128 --only $ONLY 2>&1 | tail -n +3
done > gpurun_out/mb_policy.txt 2>&1
cat gpurun_out/mb_policy.txt
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" 2>&1 | tail -5 )
timeout 120 python tools/microbench_ops.py --batch 128 --only attn_ 2>&1 | tail -3
