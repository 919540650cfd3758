#!/usr/bin/env bash
# round-1 final profile set: launch list of one training step + ncu --set full of the top kernels, summarised ON the box
# (the .ncu-rep files together exceed the 64 MiB that travel back; two small ones are kept for source-level reading)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b64.csv python tools/profile_step.py --batch 64 > /dev/null 2>&1
NCU="timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on"
: > gpurun_out/ncu_summary_s5.txt
: > gpurun_out/ncu_hotspots_s5.txt
for op in d8_qkv_headmajor d8_fc1, d8_fc2_resid dense_qkv dense_fc1_gelu dense_fc2_resid dense_fc2_dgrad_geluBwd dense_fc1_wgrad d8_fc1_wgrad attn_fwd attn_bwd gelu_d8_fwd gelu_d8_bwd ln_d8_fwd ln_d8_bwd; do
  name=${op%,}
  $NCU -o /tmp/ncu/$name -f python tools/microbench_ops.py --batch 128 --profile --only $op > /dev/null 2>&1
  echo "#### $name" >> gpurun_out/ncu_summary_s5.txt
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep >> gpurun_out/ncu_summary_s5.txt 2>&1
  echo "#### $name" >> gpurun_out/ncu_hotspots_s5.txt
  python tools/ncu_hotspots.py /tmp/ncu/$name.ncu-rep 14 >> gpurun_out/ncu_hotspots_s5.txt 2>&1
done
cp /tmp/ncu/dense_qkv.ncu-rep /tmp/ncu/d8_fc2_resid.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out/ | head -30
