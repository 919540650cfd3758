#!/usr/bin/env bash
# attention bring-up: parity tests (each param set in its own process, bounded), then timing tc vs legacy
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "test_attention" 2>&1 | tail -30 > gpurun_out/attn_tests.txt
cat gpurun_out/attn_tests.txt | tail -15
timeout 120 python tools/microbench_ops.py --batch 128 --only attn_fwd,attn_bwd 2>&1 | tail -6 | tee gpurun_out/attn_bench_tc.txt
OCTIC_ATTN_LEGACY=1 timeout 120 python tools/microbench_ops.py --batch 128 --only attn_fwd,attn_bwd 2>&1 | tail -6 | tee gpurun_out/attn_bench_legacy.txt
