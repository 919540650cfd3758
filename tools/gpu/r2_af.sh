#!/usr/bin/env bash
# round 2, call AF: ncu --set full of the octic GEMM family and the staged attention backward at the headline shapes (current code)
set -u
mkdir -p gpurun_out /tmp/ncu
NCU="timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on"
: > gpurun_out/r2af_ncu_summary.txt
: > gpurun_out/r2af_ncu_hotspots.txt
for op in d8_fc1, d8_fc2_resid d8_proj_resid d8_fc1_dgrad d8_fc2_dgrad d8_fc1_wgrad attn_bwd gelu_d8_bwd; do
  name=${op%,}
  $NCU -o /tmp/ncu/$name -f python tools/microbench_ops.py --batch 128 --profile --only $op > /dev/null 2>&1
  echo "#### $name" >> gpurun_out/r2af_ncu_summary.txt
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep >> gpurun_out/r2af_ncu_summary.txt 2>&1
  echo "#### $name" >> gpurun_out/r2af_ncu_hotspots.txt
  python tools/ncu_hotspots.py /tmp/ncu/$name.ncu-rep 14 >> gpurun_out/r2af_ncu_hotspots.txt 2>&1
done
wc -l gpurun_out/r2af_ncu_summary.txt gpurun_out/r2af_ncu_hotspots.txt
timeout 100 python tools/microbench_ops.py --batch 128 2>&1 | tail -40 > gpurun_out/r2af_microbench.txt; cat gpurun_out/r2af_microbench.txt
