#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/trace_tc2.py 1 128 2048 64 64 > gpurun_out/r2_trace_c64_d.txt 2>&1
timeout 120 python tools/trace_tc2.py 16 128 128 128 128 > gpurun_out/r2_trace_c128_d.txt 2>&1
timeout 120 python tools/trace_tc2.py 16 32 32 512 512 > gpurun_out/r2_trace_c512_d.txt 2>&1
tail -n 5 gpurun_out/r2_trace_c64_d.txt | cut -c1-400; tail -n 5 gpurun_out/r2_trace_c128_d.txt | cut -c1-600
