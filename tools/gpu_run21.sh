#!/bin/bash
mkdir -p gpurun_out
for s in layer1.0.conv1 layer3.0.ds; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 1 -f -o gpurun_out/prof_r2_simt_$s python tools/bench_small_convs.py $s > /dev/null 2>&1
done
ls -la gpurun_out/prof_r2_simt_*
