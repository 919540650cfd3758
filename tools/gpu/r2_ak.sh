#!/usr/bin/env bash
# round 2, call AK (2 GPUs): headline bench at N = 1 and N = 2 with the final kernels (staged attention backward, one-pass forward)
set -u
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2ak_bench_1gpu.json 2> gpurun_out/r2ak_bench_1gpu.err; echo "bench N=1 rc=$?"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2ak_bench_2gpu.json 2> gpurun_out/r2ak_bench_2gpu.err; echo "bench N=2 rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r2ak_bench_1gpu.json", "gpurun_out/r2ak_bench_2gpu.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d["clocks"], d["config"].get("allreduce_overlap"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/r2ak_bench_2gpu.err
