#!/usr/bin/env bash
# 8-byte store units in the attention scatter: parity, microbench, full tests, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 | cut -c1-200 )
timeout 120 python tools/microbench_ops.py --batch 128 --only attn_ 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-330 gpurun_out/bench.json
