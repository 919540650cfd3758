#!/usr/bin/env bash
# round 2, call BE: persistent staged attention backward (one CTA per SM walks its items, next operands requested under the last flushes)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention > gpurun_out/r2be_tests_attention.log 2>&1; echo "attention tests rc=$?"; tail -4 gpurun_out/r2be_tests_attention.log
for o in 1 0; do for st in 1 0; do timeout 60 build/attn_time 128 b $o $st | head -1; done; done
timeout 60 build/attn_time 192 b 1 1 | head -1
