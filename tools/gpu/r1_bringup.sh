#!/usr/bin/env bash
# Runs each GPU test in its own process (a device trap poisons the CUDA context) with a timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
FILE=${1:-tests/test_gpu_kernels.py}
TESTS=$(python -m pytest "$FILE" --collect-only -q -m gpu 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u)
for t in $TESTS; do
  echo "=== $t"
  timeout 600 python -m pytest "$t" -q -m gpu -x 2>&1 | tail -${TAILN:-25}
done 2>&1 | tee gpurun_out/bringup.log
