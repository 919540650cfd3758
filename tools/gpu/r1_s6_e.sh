#!/usr/bin/env bash
# session 6, call E (last GPU seconds): the whole GPU suite on the final tree.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 70 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/s6e_tests_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s6e_tests_gpu.log
