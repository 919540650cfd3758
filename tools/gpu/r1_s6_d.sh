#!/usr/bin/env bash
# session 6, call D (1 GPU, last GPU minutes): BASELINE configs[2]/[3] datapoints, then ncu DRAM traffic of the optimizer kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 75 python tools/bench_configs.py --steps 4 --only 2,3 > gpurun_out/s6d_bench_configs.txt 2>&1; echo "configs rc=$?"
grep config gpurun_out/s6d_bench_configs.txt | cut -c1-300
timeout 45 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:optim --csv --log-file gpurun_out/s6d_ncu_optim.csv python tools/microbench_optim.py --ncu > gpurun_out/s6d_ncu_optim.log 2>&1; echo "ncu rc=$?"
tail -12 gpurun_out/s6d_ncu_optim.csv | cut -c1-200
