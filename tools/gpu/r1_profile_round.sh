#!/usr/bin/env bash
# one GPU call: correctness after kernel edits, event table, ncu full captures of the hot kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention or gemm or linear_d8 or layerscale" 2>&1 | tail -4
python tools/profile_step.py --batch 64 --events 2>&1 | tail -26
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on"
$NCU -k regex:attn_fwd -c 1 -o gpurun_out/attn_fwd -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
$NCU -k regex:attn_bwd_kv -c 1 -o gpurun_out/attn_bwd_kv -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
$NCU -k regex:gemm_tn -s 65 -c 4 -o gpurun_out/gemm_dense -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
$NCU -k regex:gemm_tn -s 1 -c 4 -o gpurun_out/gemm_octic -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
