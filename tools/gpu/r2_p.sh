#!/usr/bin/env bash
# round 2, call P (8 GPUs): multi-bucket overlapped gradient exchange
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2p_bench_8gpu.out 2> gpurun_out/r2p_bench_8gpu.err
echo "rc=$?"
grep -h "^{" gpurun_out/r2p_bench_8gpu.out | tail -1 > gpurun_out/r2p_bench_8gpu.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r2p_bench_8gpu.json").read().strip())
print(round(d["value"],1), round(d["ms_per_step"],2), d["config"]["allreduce_overlap"], round(d["e2e"]["value"],1))
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/check_overlap_2gpu.py 2>&1 | tail -4
