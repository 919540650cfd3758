#!/usr/bin/env bash
# round 2, call AY (2 GPUs): the driver's exact launch at N = 2 with default flags (192 images per GPU), both arms
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2ay_bench_2gpu.json 2> gpurun_out/r2ay_bench_2gpu.err; echo "ours N=2 rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2ay_bench_2gpu.json') if l.startswith('{')][-1])
print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks'], d['config']['batch_per_gpu'], d['config']['allreduce_overlap'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2ay_bench_ref_2gpu.json 2> gpurun_out/r2ay_bench_ref_2gpu.err; echo "reference N=2 rc=$?"; cut -c1-200 gpurun_out/r2ay_bench_ref_2gpu.json
