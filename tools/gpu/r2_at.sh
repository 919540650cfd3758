#!/usr/bin/env bash
# round 2, call AT: default bench (192 images per GPU) after the allocator fix; 256 must fit now; 128 for the same-box ratio
set -u
mkdir -p gpurun_out
for b in 192 256 128; do
  timeout 400 python bench.py --batch $b --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2at_bench_b$b.json 2> gpurun_out/r2at_bench_b$b.err; echo "batch $b rc=$?"
  python -c "import json,torch;d=json.load(open('gpurun_out/r2at_bench_b$b.json'));print($b, round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d.get('peak_mem_gb'))"
done
timeout 400 python bench.py > gpurun_out/r2at_bench_default.json 2> gpurun_out/r2at_bench_default.err; echo "default rc=$?"; cut -c1-400 gpurun_out/r2at_bench_default.json
