#!/usr/bin/env bash
# round 2, call BC: attention backward timeline of CTA 0 at the full batch (B = 128: 13.8 CTAs per SM, full-chip contention) vs B = 20
set -u
mkdir -p gpurun_out
timeout 60 build/attn_trace 128 b 1 1 > gpurun_out/r2bc_trace_bwd_b128.txt 2>&1; echo rc=$?
timeout 60 build/attn_trace 20 b 1 1 > gpurun_out/r2bc_trace_bwd_b20.txt 2>&1; echo rc=$?
head -3 gpurun_out/r2bc_trace_bwd_b128.txt; grep -A8 "slot 1" gpurun_out/r2bc_trace_bwd_b128.txt; grep -A8 "slot 1" gpurun_out/r2bc_trace_bwd_b20.txt | tail -3
nvidia-smi --query-gpu=clocks.sm --format=csv,noheader
