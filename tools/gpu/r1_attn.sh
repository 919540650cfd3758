#!/usr/bin/env bash
# attention bring-up: parity tests (bounded), in-kernel timeline, then timing tcgen05 (head-major) vs mma.sync (packed)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "test_attention or head_remap" 2>&1 | tail -30 > gpurun_out/attn_tests.txt
tail -15 gpurun_out/attn_tests.txt
[ -x build/attn_trace ] && timeout 60 build/attn_trace 20 > gpurun_out/trace_fwd.txt 2>&1
timeout 120 python tools/microbench_ops.py --batch 128 --only attn_fwd,attn_bwd --attn-layout 2 2>&1 | tail -4 | tee gpurun_out/attn_bench_tc.txt
timeout 120 python tools/microbench_ops.py --batch 128 --only attn_fwd,attn_bwd --attn-layout 0 2>&1 | tail -4 | tee gpurun_out/attn_bench_tc_dense.txt
[ -x build/attn_trace ] && timeout 60 build/attn_trace 9 b > gpurun_out/trace_bwd.txt 2>&1
