#!/usr/bin/env bash
# round 2, call AN (8 GPUs): headline bench at N = 8 with the final kernels
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2an_bench_8gpu.json 2> gpurun_out/r2an_bench_8gpu.err; echo "bench N=8 rc=$?"
python - <<PY
import json
f = "gpurun_out/r2an_bench_8gpu.json"
try:
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d["clocks"], d["config"].get("allreduce_overlap"))
except Exception as e:
    print("unreadable", e)
PY
tail -2 gpurun_out/r2an_bench_8gpu.err | cut -c1-300
