#!/usr/bin/env bash
# ncu --set full of the tcgen05 attention kernels at the headline shape (microbench, one profiled call each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=${1:-64}
NCU="timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on"
$NCU -k regex:attn_fwd_tc -c 1 -o gpurun_out/attn_fwd_tc -f python tools/microbench_ops.py --batch $B --only attn_fwd --profile > /dev/null 2>&1
$NCU -k regex:attn_bwd_tc -c 1 -o gpurun_out/attn_bwd_tc -f python tools/microbench_ops.py --batch $B --only attn_bwd --profile > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
