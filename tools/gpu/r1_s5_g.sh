#!/usr/bin/env bash
# attention backward with deferred tile stores: parity, timeline, microbench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -k "attention or block or model" 2>&1 | tail -5 )
timeout 120 python tools/microbench_ops.py --batch 128 --only attn_ 2>&1 | tail -3
timeout 120 python tools/microbench_ops.py --batch 128 --only attn_ --attn-layout 0 2>&1 | tail -3
[ -x build/attn_trace ] && timeout 60 build/attn_trace 9 b 1 > gpurun_out/trace_bwd.txt 2>&1
head -c 3000 gpurun_out/trace_bwd.txt
