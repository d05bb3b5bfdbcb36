#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_small_m -s 25 -c 1 -f -o gpurun_out/prof_r2_linear_64_512_1536 python tools/bench_linear.py > /dev/null 2>&1
ls -la gpurun_out/prof_r2_linear_64_512_1536.ncu-rep
