#!/usr/bin/env bash
# round 2, call H: tail-column hand-off cost, TMEM read-back vs shared-memory stores
set -u
mkdir -p gpurun_out
for v in attn_trace attn_trace_nots attn_trace_nold attn_trace_nosts; do
  timeout 60 build/$v 128 b 1 > gpurun_out/r2h_${v}_b128.txt 2>&1; echo "$v: $(head -1 gpurun_out/r2h_${v}_b128.txt)"
done
