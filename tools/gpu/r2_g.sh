#!/usr/bin/env bash
# round 2, call G: what makes the tail-column hand-off expensive? (stores vs proxy fence) -- timing variants of attn_bwd
set -u
mkdir -p gpurun_out
for v in attn_trace attn_trace_nots attn_trace_nofence attn_trace_ownerfence attn_trace_nots_nofence; do
  timeout 60 build/$v 128 b 1 > gpurun_out/r2g_${v}_b128.txt 2>&1; echo "$v: $(head -1 gpurun_out/r2g_${v}_b128.txt)"
  timeout 60 build/$v 128 b 1 | head -1
done
