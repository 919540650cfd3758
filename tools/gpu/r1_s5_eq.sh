#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "equivariance_report" -s 2>&1 | tail -8 | cut -c1-1500 )
cat gpurun_out/equivariance_h14.json
