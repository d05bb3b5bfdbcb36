#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_linear_ks_sweep.txt
for ks in auto 1 2 4 8; do
  if [ $ks = auto ]; then timeout 200 python tools/bench_linear.py >> gpurun_out/r2_linear_ks_sweep.txt 2>&1; else MN_LIN_KS=$ks timeout 200 python tools/bench_linear.py >> gpurun_out/r2_linear_ks_sweep.txt 2>&1; fi
done
cat gpurun_out/r2_linear_ks_sweep.txt
