#!/usr/bin/env bash
# session 6, call F (last GPU seconds): graphed-step tests on the final tree, then the default bench line if time allows.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_model.py tests/test_gpu_optim.py -x -q -p no:cacheprovider -k "graphed" > gpurun_out/s6f_tests.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/s6f_tests.log
timeout 45 python bench.py --no-cpu-baseline > gpurun_out/s6f_bench.json 2> gpurun_out/s6f_bench.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/s6f_bench.json
