#!/usr/bin/env bash
# round 2, call R: warp-role order (tensor-feeding warps on the highest warp ids) -- correctness + A/B timing
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r2r_tests_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -2 gpurun_out/r2r_tests_kernels.log
for v in attn_time_low attn_time; do for d in f b; do echo "$v: $(timeout 60 build/$v 128 $d 1 | head -1)"; done; done
OPS=dense_qkv,dense_fc1_gelu,dense_fc2_resid,dense_fc2_dgrad_geluBwd,dense_fc1_dgrad,dense_fc1_wgrad,dense_fc2_wgrad,dense_proj_resid,d8_qkv_headmajor,d8_fc1,d8_fc2_resid,d8_proj_resid,d8_fc1_dgrad,d8_fc2_dgrad,d8_fc1_wgrad
echo "== roles high (default)"; timeout 200 python tools/microbench_ops.py --batch 128 --only $OPS 2>&1 | tail -16 | tee gpurun_out/r2r_microbench_roles_high.txt
echo "== roles low (round 1)"; OCTIC_GEMM_ROLES_LOW=1 timeout 200 python tools/microbench_ops.py --batch 128 --only $OPS 2>&1 | tail -16 | tee gpurun_out/r2r_microbench_roles_low.txt
