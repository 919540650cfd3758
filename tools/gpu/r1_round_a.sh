#!/usr/bin/env bash
# one GPU call: tests, bench, event table, launch list, ncu full captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
( time python -m pytest tests -q -m gpu -x 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log 2>&1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python tools/profile_step.py --batch 128 --events > gpurun_out/events_b128.txt 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b64.csv python tools/profile_step.py --batch 64 > /dev/null 2>&1
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on"
$NCU -k regex:attn_fwd -c 1 -o gpurun_out/attn_fwd -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
$NCU -k regex:attn_bwd_kv -c 1 -o gpurun_out/attn_bwd_kv -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
$NCU -k regex:attn_bwd_q -c 1 -o gpurun_out/attn_bwd_q -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
$NCU -k regex:gemm_tn -s 70 -c 4 -o gpurun_out/gemm_dense -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
$NCU -k regex:gemm_tn -s 1 -c 4 -o gpurun_out/gemm_octic -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
$NCU -k regex:gemm_wgrad -s 4 -c 3 -o gpurun_out/gemm_wgrad -f python tools/profile_step.py --batch 32 > /dev/null 2>&1
ls -la gpurun_out/
cat gpurun_out/pytest_gpu.log gpurun_out/bench.json; tail -3 gpurun_out/bench.err
