#!/usr/bin/env bash
# round 2, call Q: ncu --set full of the kernels that are furthest from their roofline (one launch each at the headline shapes)
set -u
mkdir -p gpurun_out /tmp/ncu
NCU="timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on"
: > gpurun_out/r2q_ncu_summary.txt
: > gpurun_out/r2q_ncu_hotspots.txt
for op in ln_d8_fwd ln_d8_bwd ln_fwd ln_bwd gelu_d8_fwd gelu_d8_bwd d8_fc1, d8_qkv_headmajor dense_fc1_wgrad dense_fc1_gelu attn_fwd attn_bwd; do
  name=${op%,}
  $NCU -o /tmp/ncu/$name -f python tools/microbench_ops.py --batch 128 --profile --only $op > /dev/null 2>&1
  echo "#### $name" >> gpurun_out/r2q_ncu_summary.txt
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep >> gpurun_out/r2q_ncu_summary.txt 2>&1
  echo "#### $name" >> gpurun_out/r2q_ncu_hotspots.txt
  python tools/ncu_hotspots.py /tmp/ncu/$name.ncu-rep 12 >> gpurun_out/r2q_ncu_hotspots.txt 2>&1
done
wc -l gpurun_out/r2q_ncu_summary.txt gpurun_out/r2q_ncu_hotspots.txt
