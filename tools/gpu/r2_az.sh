#!/usr/bin/env bash
# round 2, call AZ: BASELINE.json configs[1]-[3] datapoints with the round-2 kernels; DINOv2 break-layer sweep; ncu of the new attention forward
set -u
mkdir -p gpurun_out /tmp/ncu
timeout 600 python tools/bench_configs.py --steps 6 > gpurun_out/r2az_bench_configs.txt 2>&1; echo "configs rc=$?"; grep -v Warning gpurun_out/r2az_bench_configs.txt | tail -5 | cut -c1-400
timeout 600 python tools/bench_dinov2.py --batch 64 --steps 4 --layers 0,12,23 > gpurun_out/r2az_bench_dinov2.txt 2>&1; echo "dinov2 rc=$?"; grep octic_equi gpurun_out/r2az_bench_dinov2.txt | cut -c1-200
NCU="timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on"
: > gpurun_out/r2az_ncu_summary.txt
for op in attn_fwd dense_fc1_gelu; do
  $NCU -o /tmp/ncu/$op -f python tools/microbench_ops.py --batch 128 --profile --only $op > /dev/null 2>&1
  echo "#### $op" >> gpurun_out/r2az_ncu_summary.txt
  python tools/ncu_summary.py /tmp/ncu/$op.ncu-rep >> gpurun_out/r2az_ncu_summary.txt 2>&1
done
grep -E "^####|^==|time_duration|tensor_cycles_active|issue_active|dram__bytes|xu_cycles" gpurun_out/r2az_ncu_summary.txt | cut -c1-150
