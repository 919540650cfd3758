#!/usr/bin/env bash
# round 2, call U: N-tile width of the LinearD8 launches
set -u
OPS=d8_qkv,d8_fc1,d8_fc2_resid,d8_proj_resid,d8_fc1_dgrad,d8_fc2_dgrad,d8_qkv_headmajor
for bn in 0 64 96 128 192 256; do
  echo "== OCTIC_BLOCK_N=$bn"; OCTIC_BLOCK_N=$bn timeout 200 python tools/microbench_ops.py --batch 128 --only $OPS 2>&1 | grep -E "^d8_" | awk '{printf "%s %s | ", $1, $2} END {print ""}'
done
timeout 100 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear_d8" 2>&1 | tail -1
OCTIC_BLOCK_N=128 timeout 100 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear_d8" 2>&1 | tail -1
OCTIC_BLOCK_N=256 timeout 100 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear_d8" 2>&1 | tail -1
