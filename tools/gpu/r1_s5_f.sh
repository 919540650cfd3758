#!/usr/bin/env bash
# policy check (staged epilogues, pairs by shape): full tests, events, bench; attention in-kernel timelines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 ) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
cat gpurun_out/events_b128.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
[ -x build/attn_trace ] && timeout 60 build/attn_trace 9 b 1 > gpurun_out/trace_bwd.txt 2>&1
[ -x build/attn_trace ] && timeout 60 build/attn_trace 18 f 1 > gpurun_out/trace_fwd.txt 2>&1
head -c 6000 gpurun_out/trace_bwd.txt
