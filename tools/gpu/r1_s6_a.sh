#!/usr/bin/env bash
# session 6, call A: new DINOv2 + fused-optimizer GPU tests, tests touched by the FlatGrads change, optimizer
# microbench (GB/s vs HBM peak), one bench line with the fused LAMB step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dinov2.py tests/test_gpu_optim.py -q --tb=short -p no:cacheprovider > gpurun_out/s6a_tests_new.log 2>&1
tail -40 gpurun_out/s6a_tests_new.log
timeout 300 python -m pytest tests/test_gpu_model.py -q --tb=short -p no:cacheprovider -k "graphed or whole_model or error or timm" > gpurun_out/s6a_tests_model.log 2>&1
tail -5 gpurun_out/s6a_tests_model.log
timeout 200 python tools/microbench_optim.py > gpurun_out/s6a_optim_microbench.txt 2>&1
cat gpurun_out/s6a_optim_microbench.txt
timeout 240 python tools/bench_dinov2.py --batch 64 --steps 4 --layers 0,12,23 > gpurun_out/s6a_bench_dinov2.txt 2>&1
tail -5 gpurun_out/s6a_bench_dinov2.txt
timeout 300 python bench.py --steps 6 --warmup 3 --optimizer lamb --no-cpu-baseline > gpurun_out/s6a_bench_lamb.json 2> gpurun_out/s6a_bench_lamb.err
tail -3 gpurun_out/s6a_bench_lamb.err; cat gpurun_out/s6a_bench_lamb.json
