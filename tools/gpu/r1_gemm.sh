#!/usr/bin/env bash
# GEMM kernels: parity tests + stand-alone timing at the headline shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "gemm or linear or dense" 2>&1 | tail -4
timeout 300 python tools/microbench_ops.py --batch 128 --only d8_,dense_ 2>&1 | tail -14 | tee gpurun_out/gemm_bench.txt
