#!/usr/bin/env bash
# round 2, call W: lane-interleaved LayerNormD8 forward + tuned LinearD8 launch policy; tests + microbench + bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r2w_tests_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -2 gpurun_out/r2w_tests_kernels.log
echo "== il"; timeout 100 python tools/microbench_ops.py --batch 128 --only ln_ 2>&1 | grep -E "^ln_"
echo "== lc"; OCTIC_LN_LC=1 timeout 100 python tools/microbench_ops.py --batch 128 --only ln_ 2>&1 | grep -E "^ln_"
timeout 100 python tools/microbench_ops.py --batch 128 --only d8_ 2>&1 | grep -E "^d8_"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2w_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
print(d["step_breakdown_ms"])
PY
