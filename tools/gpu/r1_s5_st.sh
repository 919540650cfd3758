#!/usr/bin/env bash
# TMA ring depth sensitivity of the octic / dense GEMMs
cd "$(dirname "$0")/.."
for st in 0 3 2; do
  echo "## OCTIC_GEMM_STAGES=$st (0 = as many as fit)"
  OCTIC_GEMM_STAGES=$st timeout 200 python tools/microbench_ops.py --batch 128 --iters 20 --only d8_qkv,d8_proj_resid,d8_fc1,d8_fc2,dense_qkv,dense_proj_plain 2>&1 | tail -n +3 | grep -v wgrad
done
