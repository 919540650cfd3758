#!/usr/bin/env bash
# round 2, call Y: delta fused into the attention backward prologue; event table
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention > gpurun_out/r2y_tests_attention.log 2>&1; echo "attention tests rc=$?"; tail -2 gpurun_out/r2y_tests_attention.log
for o in 1 0; do timeout 60 build/attn_time 128 b $o | head -1; done
timeout 100 python tools/microbench_ops.py --batch 128 --only attn 2>&1 | grep -E "^attn"
timeout 200 python tools/profile_step.py --batch 128 --events > gpurun_out/r2y_events_b128.txt 2>&1; echo "events rc=$?"
cat gpurun_out/r2y_events_b128.txt | tail -26
