#!/usr/bin/env bash
# round 2, call V: LinearD8 launches with the new tile rule -- pairs and ring depth
set -u
OPS=d8_qkv,d8_fc1,d8_fc2_resid,d8_proj_resid,d8_fc1_dgrad,d8_fc2_dgrad,d8_qkv_headmajor
run() { echo "== $1"; env $1 timeout 200 python tools/microbench_ops.py --batch 128 --only $OPS 2>&1 | grep -E "^d8_" | awk '{printf "%s %s | ", $1, $2} END {print ""}'; }
run "X=default"
run "OCTIC_GEMM_NCTA=2"
run "OCTIC_GEMM_NCTA=1"
for st in 2 3 4 6; do run "OCTIC_GEMM_STAGES=$st"; done
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear_d8 or gemm" 2>&1 | tail -1
