#!/usr/bin/env bash
# round 2, call BB: operand loads issued in the prologue (before the scratch-slot claim / TMEM allocation): tests + timing
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention > gpurun_out/r2bb_tests_attention.log 2>&1; echo "attention tests rc=$?"; tail -3 gpurun_out/r2bb_tests_attention.log
for o in 1 0; do timeout 60 build/attn_time 128 b $o 1 | head -1; timeout 60 build/attn_time 128 f $o | head -1; done
