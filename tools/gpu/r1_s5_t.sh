#!/usr/bin/env bash
# delta kernel (warp per row), smoke(), batch-size sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 | cut -c1-200 )
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | cut -c1-600
timeout 120 python tools/microbench_ops.py --batch 128 --only attn_bwd,gelu_d8 2>&1 | tail -3
for b in 128 148 192; do
  timeout 600 python bench.py --no-cpu-baseline --batch $b > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_b$b.json'));print($b, round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
done
