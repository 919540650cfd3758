#!/usr/bin/env bash
# round-1 profile set: launch list of one training step + ncu --set full of the top kernels (via the microbench harness)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b64.csv python tools/profile_step.py --batch 64 > /dev/null 2>&1
NCU="timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on"
$NCU -o gpurun_out/ncu_d8_qkv_headmajor -f python tools/microbench_ops.py --batch 128 --profile --only d8_qkv_headmajor > /dev/null 2>&1
$NCU -o gpurun_out/ncu_dense_qkv -f python tools/microbench_ops.py --batch 128 --profile --only dense_qkv > /dev/null 2>&1
$NCU -k regex:gemm_wgrad -o gpurun_out/ncu_dense_fc1_wgrad -f python tools/microbench_ops.py --batch 128 --profile --only dense_fc1_wgrad > /dev/null 2>&1
$NCU -o gpurun_out/ncu_dense_fc1_gelu -f python tools/microbench_ops.py --batch 128 --profile --only dense_fc1_gelu > /dev/null 2>&1
$NCU -k regex:attn_bwd_tc -o gpurun_out/ncu_attn_bwd -f python tools/microbench_ops.py --batch 128 --profile --only attn_bwd > /dev/null 2>&1
$NCU -k regex:attn_fwd_tc -o gpurun_out/ncu_attn_fwd -f python tools/microbench_ops.py --batch 128 --profile --only attn_fwd > /dev/null 2>&1
$NCU -o gpurun_out/ncu_d8_fc1 -f python tools/microbench_ops.py --batch 128 --profile --only d8_fc1, > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_b64.csv
