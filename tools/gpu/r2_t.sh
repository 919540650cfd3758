#!/usr/bin/env bash
set -u
OPS=d8_qkv,d8_fc1,d8_fc2_resid,d8_proj_resid,d8_fc1_dgrad,d8_fc2_dgrad,dense_proj_plain,dense_qkv
echo "== packed (default)"; timeout 200 python tools/microbench_ops.py --batch 128 --only $OPS 2>&1 | grep -E "^d8_|^dense" | tee gpurun_out/r2t_microbench_packed.txt
echo "== OCTIC_GEMM_PACKED=0"; OCTIC_GEMM_PACKED=0 timeout 200 python tools/microbench_ops.py --batch 128 --only $OPS 2>&1 | grep -E "^d8_|^dense" | tee gpurun_out/r2t_microbench_fp32staged.txt
