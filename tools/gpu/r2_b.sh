#!/usr/bin/env bash
# round 2, call B: epilogue probe, in-kernel attention timelines, today's per-op event table, attention tests after the xch fix.
set -u
mkdir -p gpurun_out
{
  for v in 0 1 2 3; do for w in 4 8 16; do timeout 30 build/epilogue_probe $v $w 5 || echo "variant $v warps $w failed (rc=$?)"; done; done
} > gpurun_out/r2b_epilogue_probe.txt 2>&1
tail -15 gpurun_out/r2b_epilogue_probe.txt
timeout 60 build/attn_trace 20 f 1 > gpurun_out/r2b_trace_fwd_octic.txt 2>&1; echo "trace fwd rc=$?"
timeout 60 build/attn_trace 20 b 1 > gpurun_out/r2b_trace_bwd_octic.txt 2>&1; echo "trace bwd rc=$?"
timeout 60 build/attn_trace 20 b 0 > gpurun_out/r2b_trace_bwd_dense.txt 2>&1; echo "trace bwd dense rc=$?"
timeout 120 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention > gpurun_out/r2b_tests_attention.log 2>&1; echo "attention tests rc=$?"
tail -2 gpurun_out/r2b_tests_attention.log
timeout 200 python tools/profile_step.py --batch 128 --events > gpurun_out/r2b_events_b128.txt 2>&1; echo "events rc=$?"
tail -30 gpurun_out/r2b_events_b128.txt
