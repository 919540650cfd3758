#!/usr/bin/env bash
# session-4 first call: full GPU tests, bench, event table, attention tc-vs-legacy timing, launch list, ncu full on attention tc
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
timeout 120 python tools/microbench_ops.py --batch 128 --only attn_fwd,attn_bwd 2>&1 | tail -6 > gpurun_out/attn_bench_tc.txt
OCTIC_ATTN_LEGACY=1 timeout 120 python tools/microbench_ops.py --batch 128 --only attn_fwd,attn_bwd 2>&1 | tail -6 > gpurun_out/attn_bench_legacy.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b64.csv python tools/profile_step.py --batch 64 > /dev/null 2>&1
NCU="timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on"
$NCU -k regex:attn_fwd_tc -c 1 -o gpurun_out/attn_fwd_tc -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
$NCU -k regex:attn_bwd_tc -c 1 -o gpurun_out/attn_bwd_tc -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
ls -la gpurun_out/
cat gpurun_out/pytest_gpu.log gpurun_out/bench.json gpurun_out/events_b128.txt gpurun_out/attn_bench_tc.txt gpurun_out/attn_bench_legacy.txt; tail -3 gpurun_out/bench.err
