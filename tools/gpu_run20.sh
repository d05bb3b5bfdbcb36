#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_small_convs.py > gpurun_out/r2_small_convs.txt 2>&1
cat gpurun_out/r2_small_convs.txt | cut -c1-200
