#!/usr/bin/env bash
# Round-2 opener (about one GPU-minute): which epilogue scheme can drain 128x256 accumulator tiles at HBM speed?
# Decides the redesign of the octic (small-K) GEMM epilogue -- see tools/epilogue_probe.cu and DESIGN.md section 3.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/epilogue_probe tools/epilogue_probe.cu || exit 1
{
  for v in 0 1 2 3; do
    for w in 4 8 16; do
      timeout 30 build/epilogue_probe $v $w 5 || echo "variant $v warps $w failed (rc=$?)"
    done
  done
} 2>&1 | tee gpurun_out/epilogue_probe.txt
