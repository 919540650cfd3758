#!/usr/bin/env bash
# round 2, call O (8 GPUs): scaling of the default (overlapped) gradient exchange vs one all-reduce after the replay.
set -u
mkdir -p gpurun_out
for mode in "" "--no-overlap"; do
  tag=$([ -z "$mode" ] && echo overlap || echo plain)
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 10 --warmup 3 $mode > gpurun_out/r2o_bench_8gpu_$tag.out 2> gpurun_out/r2o_bench_8gpu_$tag.err
  echo "$tag rc=$?"
  grep -h "^{" gpurun_out/r2o_bench_8gpu_$tag.out | tail -1 > gpurun_out/r2o_bench_8gpu_$tag.json
  grep -c "NVLS" gpurun_out/r2o_bench_8gpu_$tag.out gpurun_out/r2o_bench_8gpu_$tag.err | head -2
  grep -h "NVLS\|Channel.*nvls\|algorithm" gpurun_out/r2o_bench_8gpu_$tag.out gpurun_out/r2o_bench_8gpu_$tag.err | head -4 | cut -c1-200
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2o_bench_8gpu_$tag.json").read().strip())
    print("$tag", round(d["value"],1), round(d["ms_per_step"],2), d["config"]["allreduce_overlap"], round(d["e2e"]["value"],1))
except Exception as e:
    print("no json:", e)
PY
done
