#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_linear_ks_sweep_b.txt
for ks in auto 1 2 4; do
  if [ $ks = auto ]; then timeout 200 python tools/bench_linear.py >> gpurun_out/r2_linear_ks_sweep_b.txt 2>&1; else MN_LIN_KS=$ks timeout 200 python tools/bench_linear.py >> gpurun_out/r2_linear_ks_sweep_b.txt 2>&1; fi
done
cat gpurun_out/r2_linear_ks_sweep_b.txt
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
