#!/usr/bin/env bash
# round 2, call AQ: the reference arm of bench.py on the GPU box's host cores (what the driver runs first at round end)
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2aq_bench_reference.json 2> gpurun_out/r2aq_bench_reference.err; echo "reference arm rc=$?"
cat gpurun_out/r2aq_bench_reference.json | cut -c1-900; tail -2 gpurun_out/r2aq_bench_reference.err | cut -c1-300
