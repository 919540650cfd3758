#!/usr/bin/env bash
# round 2, call AC: staged dQ + packed-pair math + 128-byte scratch pitch: attention tests, timing, timeline, model tests, bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention > gpurun_out/r2ac_tests_attention.log 2>&1; echo "attention tests rc=$?"; tail -3 gpurun_out/r2ac_tests_attention.log
for o in 1 0; do for st in 1 0; do timeout 60 build/attn_time 128 b $o $st | head -1; done; done
timeout 60 build/attn_trace 20 b 1 1 > gpurun_out/r2ac_trace_bwd_staged.txt 2>&1; echo "trace rc=$?"
head -16 gpurun_out/r2ac_trace_bwd_staged.txt
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_budget.py -x -q -m gpu > gpurun_out/r2ac_tests_model.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2ac_tests_model.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err; echo "bench rc=$?"; cat gpurun_out/r2ac_bench.json | head -c 1500
