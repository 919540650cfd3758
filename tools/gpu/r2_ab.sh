#!/usr/bin/env bash
# round 2, call AB: staged dQ in the tcgen05 attention backward (dS^T through an L2-resident scratch): tests, timing, timeline
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention > gpurun_out/r2ab_tests_attention.log 2>&1; echo "attention tests rc=$?"; tail -3 gpurun_out/r2ab_tests_attention.log
for o in 1 0; do for st in 1 0; do timeout 60 build/attn_time 128 b $o $st | head -1; done; done
for br in 136 68 16; do echo "ds box rows $br"; OCTIC_DS_BOX_ROWS=$br timeout 60 build/attn_time 128 b 1 1 | head -1; done
timeout 60 build/attn_trace 20 b 1 1 > gpurun_out/r2ab_trace_bwd_staged.txt 2>&1; echo "trace rc=$?"
head -14 gpurun_out/r2ab_trace_bwd_staged.txt
OCTIC_DS_BOX_ROWS=16 timeout 60 build/attn_trace 20 b 1 1 | head -6
