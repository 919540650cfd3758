#!/usr/bin/env bash
# round 2, call BA: number of dS^T scratch slots of the staged attention backward (L2 working set) -- 296 (default) vs fewer
set -u
for s in 0 148 152 160 200 296 592; do echo "slots $s"; timeout 60 build/attn_time 128 b 1 1 $s | head -1; done
