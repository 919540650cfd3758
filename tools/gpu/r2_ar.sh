#!/usr/bin/env bash
# round 2, call AR: images per GPU 64 / 128 / 192 / 256 on one box (SURVEY 8d allows 64-256 for the headline metric)
set -u
mkdir -p gpurun_out
for b in 128 64 192 256 128; do
  timeout 300 python bench.py --batch $b --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2ar_bench_b$b.json 2> gpurun_out/r2ar_bench.err; echo "batch $b rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2ar_bench_b$b.json'));print($b, round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
done
