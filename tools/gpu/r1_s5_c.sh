#!/usr/bin/env bash
# 2-CTA GEMM + attention tail bring-up: targeted tests first (bounded), then everything, bench, events, A/B vs 1-CTA
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "gemm" 2>&1 | tail -15 ) > gpurun_out/pytest_gemm.log 2>&1
cat gpurun_out/pytest_gemm.log
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" 2>&1 | tail -15 ) > gpurun_out/pytest_attn.log 2>&1
cat gpurun_out/pytest_attn.log
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
cat gpurun_out/events_b128.txt
OCTIC_GEMM_NCTA=1 timeout 300 python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128_1cta.txt 2>&1
head -12 gpurun_out/events_b128_1cta.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
