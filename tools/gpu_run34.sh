#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_conv_gn.txt
for shape in "16 64 64 512 256" "16 64 64 256 256" "1 64 1024 256 256" "1 128 2048 64 64"; do
  timeout 100 python tools/bench_conv_gn.py $shape 10 >> gpurun_out/r2_conv_gn.txt 2>&1
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_tc2_gn_fused python tools/bench_conv_gn.py 16 64 64 512 256 3 > /dev/null 2>&1
cat gpurun_out/r2_conv_gn.txt
