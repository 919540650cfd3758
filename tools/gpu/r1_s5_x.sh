#!/usr/bin/env bash
# A/B on one box: 8-byte store units vs 4-byte units in the attention scatter
cd "$(dirname "$0")/.."
for i in 1 2; do
  for u in 1 0; do
    echo "## OCTIC_ATTN_UNITS=$u"
    OCTIC_ATTN_UNITS=$u timeout 120 python tools/microbench_ops.py --batch 128 --iters 30 --only attn_ 2>&1 | tail -2
  done
done
