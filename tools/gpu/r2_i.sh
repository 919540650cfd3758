#!/usr/bin/env bash
# round 2, call I: timelines of a bsel=0 and a bsel=1 math warp with / without the tail-column hand-off
set -u
mkdir -p gpurun_out
for v in attn_trace attn_trace_nots; do
  timeout 60 build/$v 20 b 1 > gpurun_out/r2i_${v}_b20.txt 2>&1; echo "$v: $(head -1 gpurun_out/r2i_${v}_b20.txt)"
done
