#!/usr/bin/env bash
# round 2, call AS: why does batch 256 fail; batch 224
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --batch 256 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2as_bench_b256.json 2> gpurun_out/r2as_bench_b256.err; echo "batch 256 rc=$?"
grep -v "Warning\|warn\|^  return\|^$" gpurun_out/r2as_bench_b256.err | tail -12 | cut -c1-400
timeout 300 python bench.py --batch 224 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2as_bench_b224.json 2> gpurun_out/r2as_bench_b224.err; echo "batch 224 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r2as_bench_b224.json'));print(224, round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
nvidia-smi --query-gpu=memory.total --format=csv
