#!/usr/bin/env bash
# round 2, call AL: attention backward timeline, dense layout (16-byte stores) vs octic (4-byte scatter): what the flush costs
set -u
mkdir -p gpurun_out
timeout 60 build/attn_trace 20 b 0 1 > gpurun_out/r2al_trace_bwd_dense.txt 2>&1; echo "rc=$?"
timeout 60 build/attn_trace 20 b 1 1 > gpurun_out/r2al_trace_bwd_octic.txt 2>&1; echo "rc=$?"
cat gpurun_out/r2al_trace_bwd_dense.txt; grep -A8 "slot 1" gpurun_out/r2al_trace_bwd_octic.txt
