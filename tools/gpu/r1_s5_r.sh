#!/usr/bin/env bash
# attention: both warp sets share every job -- parity, microbench, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -k "attention or block or model or golden" 2>&1 | tail -6 | cut -c1-300 )
timeout 120 python tools/microbench_ops.py --batch 128 --only attn_ 2>&1 | tail -3
timeout 120 python tools/microbench_ops.py --batch 128 --only attn_ --attn-layout 0 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-330 gpurun_out/bench.json
