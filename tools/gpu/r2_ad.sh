#!/usr/bin/env bash
# round 2, call AD: staged dQ (scalar math, 128-byte scratch pitch): timing, timeline, same-box A/B of the bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention > gpurun_out/r2ad_tests_attention.log 2>&1; echo "attention tests rc=$?"; tail -3 gpurun_out/r2ad_tests_attention.log
for o in 1 0; do for st in 1 0; do timeout 60 build/attn_time 128 b $o $st | head -1; done; done
timeout 60 build/attn_trace 20 b 1 1 > gpurun_out/r2ad_trace_bwd_staged.txt 2>&1; echo "trace rc=$?"
head -16 gpurun_out/r2ad_trace_bwd_staged.txt
for st in 1 0 1 0; do
  OCTIC_ATTN_STAGED_DQ=$st timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2ad_bench_staged$st.json 2> gpurun_out/r2ad_bench.err; echo "bench staged=$st rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2ad_bench_staged$st.json'));print(d['value'], d['ms_per_step'], d['clocks'])"
done
