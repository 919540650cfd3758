#!/usr/bin/env bash
# round 2, call BF: ncu launch list of bench.py itself (the recipe's gpu__time_duration pass), default config
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2bf_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2bf_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2bf_launches_bench.csv 40 > gpurun_out/r2bf_launches_bench.txt 2>&1; head -30 gpurun_out/r2bf_launches_bench.txt
rm -f gpurun_out/r2bf_launches_bench.csv
