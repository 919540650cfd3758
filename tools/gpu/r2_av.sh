#!/usr/bin/env bash
# round 2, call AV: final tree -- full GPU suite, smoke, default bench (ours), launch list at the bench batch
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2av_tests_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2av_tests_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2av_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2av_smoke.log
timeout 400 python bench.py > gpurun_out/r2av_bench.json 2> gpurun_out/r2av_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2av_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['gpu_launches'], d['roofline']['frac'], d['roofline_octic']['frac'], d['tc_util_vs_sustained_peak'], d['config']['batch_per_gpu'])"
