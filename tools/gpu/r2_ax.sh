#!/usr/bin/env bash
# round 2, call AX: weight-gradient GEMMs on a side stream inside the graphed step (prototype, OCTIC_SIDE_WGRAD=1): A/B on one box
set -u
mkdir -p gpurun_out
for sw in 1 0 1 0; do
  OCTIC_SIDE_WGRAD=$sw timeout 300 python bench.py --batch 128 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2ax_bench_side$sw.json 2> gpurun_out/r2ax_bench_side$sw.err; echo "side=$sw rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2ax_bench_side$sw.json'));print($sw, round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['config']['cuda_graph'])"
done
OCTIC_SIDE_WGRAD=1 timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "graphed or golden" > gpurun_out/r2ax_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ax_tests.log
