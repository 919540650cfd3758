#!/usr/bin/env bash
# after the cluster-arrive fix: GEMM tests, pairs vs single microbench, event table, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "gemm or linear" 2>&1 | tail -4 )
ONLY=d8_,dense_
for n in 0 2 1; do
  echo "## OCTIC_GEMM_NCTA=$n (0 = policy)"
  OCTIC_GEMM_NCTA=$n timeout 300 python tools/microbench_ops.py --batch 128 --only $ONLY 2>&1 | tail -n +3
done > gpurun_out/mb_policy2.txt 2>&1
cat gpurun_out/mb_policy2.txt
timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
head -12 gpurun_out/events_b128.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err | cut -c1-300
