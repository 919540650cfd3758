#!/usr/bin/env bash
# round 2, call M: new parity-budget / boundary / equivariance tests; racecheck of the attention forward after the fix.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_budget.json
timeout 900 python -m pytest tests/test_gpu_parity_budget.py tests/test_gpu_boundary.py -x -q -s -m gpu > gpurun_out/r2m_tests_new.log 2>&1; echo "new tests rc=$?"
tail -60 gpurun_out/r2m_tests_new.log
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -m gpu -k "equivariance_report or failed_graph or scoped" > gpurun_out/r2m_tests_model.log 2>&1; echo "model tests rc=$?"
tail -8 gpurun_out/r2m_tests_model.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "test_attention and 17-2-64 or test_attention and 129-2-80" > gpurun_out/r2m_racecheck_attention.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r2m_racecheck_attention.log
