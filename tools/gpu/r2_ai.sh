#!/usr/bin/env bash
# round 2, call AI: one-pass attention forward (lazy rescale): tests, timing, timeline, in-step event tables (train + inference)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention > gpurun_out/r2ai_tests_attention.log 2>&1; echo "attention tests rc=$?"; tail -4 gpurun_out/r2ai_tests_attention.log
for o in 1 0; do timeout 60 build/attn_time 128 f $o | head -1; done
timeout 60 build/attn_trace 20 f 1 > gpurun_out/r2ai_trace_fwd.txt 2>&1; echo "trace rc=$?"
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_budget.py tests/test_gpu_dinov2.py -x -q -m gpu > gpurun_out/r2ai_tests_model.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2ai_tests_model.log
timeout 200 python tools/profile_step.py --batch 128 --events > gpurun_out/r2ai_events_b128.txt 2>&1; grep -E "step total|attention" gpurun_out/r2ai_events_b128.txt
timeout 200 python tools/profile_step.py --batch 256 --infer --events > gpurun_out/r2ai_events_infer_b256.txt 2>&1; grep -E "step total|attention" gpurun_out/r2ai_events_infer_b256.txt
