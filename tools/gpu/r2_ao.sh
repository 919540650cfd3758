#!/usr/bin/env bash
# round 2, call AO: attention backward -- tail dQ rows on the (idle) tail warps, accumulator flush deferred behind the next tile's first job
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention > gpurun_out/r2ao_tests_attention.log 2>&1; echo "attention tests rc=$?"; tail -4 gpurun_out/r2ao_tests_attention.log
for o in 1 0; do for st in 1 0; do timeout 60 build/attn_time 128 b $o $st | head -1; done; done
timeout 60 build/attn_trace 20 b 1 1 > gpurun_out/r2ao_trace_bwd_octic.txt 2>&1; echo "trace rc=$?"
cat gpurun_out/r2ao_trace_bwd_octic.txt | head -16
