#!/usr/bin/env bash
# round 2, call J: attention bwd with the tail-column store out of line; parity + timing (no trace code in the timed builds)
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r2j_tests_attention.log 2>&1; echo "attention tests rc=$?"
tail -2 gpurun_out/r2j_tests_attention.log
for v in attn_time attn_time_nots; do
  for o in 1 0; do timeout 60 build/$v 128 b $o | head -1; done
done
timeout 60 build/attn_time 128 f 1 | head -1
timeout 60 build/attn_trace 20 b 1 > gpurun_out/r2j_trace_bwd_b20.txt 2>&1
timeout 120 python tools/microbench_ops.py --batch 128 --only attn > gpurun_out/r2j_microbench_attn.txt 2>&1; tail -2 gpurun_out/r2j_microbench_attn.txt
