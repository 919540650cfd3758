#!/usr/bin/env bash
# session 6, call B (1 GPU): the whole GPU suite (timed), smoke(), default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=8 > gpurun_out/s6b_tests_gpu.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"
tail -15 gpurun_out/s6b_tests_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s6b_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/s6b_smoke.log
timeout 300 python bench.py > gpurun_out/s6b_bench.json 2> gpurun_out/s6b_bench.err; echo "bench rc=$? total ${SECONDS}s"
cat gpurun_out/s6b_bench.json | cut -c1-300
