#!/usr/bin/env bash
# session 6, call C (2 GPUs): early all-reduce inside the graph (correctness + replica consistency), 2-GPU bench with
# and without the overlap.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR tools/check_overlap_2gpu.py > gpurun_out/s6c_check_overlap.log 2>&1; echo "check rc=$?"
grep -E "rank|ok|Error|error" gpurun_out/s6c_check_overlap.log | tail -12
timeout 200 $TR bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/s6c_bench_2gpu_overlap.json 2> gpurun_out/s6c_bench_2gpu_overlap.err; echo "bench overlap rc=$?"
tail -2 gpurun_out/s6c_bench_2gpu_overlap.err; cat gpurun_out/s6c_bench_2gpu_overlap.json | cut -c1-400
timeout 200 $TR bench.py --gpus 2 --steps 6 --warmup 3 --no-overlap > gpurun_out/s6c_bench_2gpu_plain.json 2> gpurun_out/s6c_bench_2gpu_plain.err; echo "bench plain rc=$?"
cat gpurun_out/s6c_bench_2gpu_plain.json | cut -c1-400
