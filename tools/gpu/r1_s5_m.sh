#!/usr/bin/env bash
# wgrad: whole tiles first + stream-K remainder, vector RED: tests, microbench, event table, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "wgrad or linear_d8 or gamma_folded" 2>&1 | tail -4 )
timeout 300 python tools/microbench_ops.py --batch 128 --only dense_fc1_wgrad,dense_fc2_wgrad,dense_qkv_wgrad,dense_proj_wgrad,d8_fc1_wgrad,d8_qkv_wgrad 2>&1 | tail -n +2 | tee gpurun_out/mb_wgrad.txt
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 ) 2>&1 | tail -8
timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
head -12 gpurun_out/events_b128.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
