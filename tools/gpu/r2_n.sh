#!/usr/bin/env bash
# round 2, call N (2 GPUs): gradient all-reduce overlap as the default -- timing vs --no-overlap, clean exit.
set -u
mkdir -p gpurun_out
for mode in "" "--no-overlap"; do
  tag=$([ -z "$mode" ] && echo overlap || echo plain)
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 $mode > gpurun_out/r2n_bench_2gpu_$tag.json 2> gpurun_out/r2n_bench_2gpu_$tag.err
  echo "$tag rc=$?"; tail -1 gpurun_out/r2n_bench_2gpu_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2n_bench_2gpu_$tag.json").read().strip().splitlines()[-1])
    print("$tag", d["value"], d["ms_per_step"], d["config"]["allreduce_overlap"], d["e2e"]["value"])
except Exception as e:
    print("no json:", e)
PY
done
echo done
