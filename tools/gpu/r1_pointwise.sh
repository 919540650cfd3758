#!/usr/bin/env bash
# pointwise kernels: parity tests, stand-alone timing at the headline shape, ncu full captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "not attention" 2>&1 | tail -5
timeout 200 python tools/microbench_ops.py --batch 128 --only ln_,gelu,layerscale,colsum 2>&1 | tail -12 | tee gpurun_out/pointwise_bench.txt
if [ "$1" = "ncu" ]; then
NCU="timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on"
for k in ln_d8_fwd ln_d8_bwd layerscale_bwd gelu_d8_fwd; do
  $NCU -c 1 -o gpurun_out/pw_$k -f python tools/microbench_ops.py --batch 128 --only $k --profile > /dev/null 2>&1
done
ls -la gpurun_out/pw_*.ncu-rep
fi
