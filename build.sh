#!/usr/bin/env bash
# Builds liboctic_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
SRC=octic_vits_b200/csrc
OUT=octic_vits_b200/lib
mkdir -p "$OUT" build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --use_fast_math -Xptxas -v $*"
pids=()
for f in gemm_sm100 pointwise attention attention_tc capi optimizer; do
  ( $NVCC $FLAGS -c $SRC/$f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT/liboctic_b200.so build/gemm_sm100.o build/pointwise.o build/attention.o build/attention_tc.o build/capi.o build/optimizer.o -lcudart
echo "built $OUT/liboctic_b200.so"
