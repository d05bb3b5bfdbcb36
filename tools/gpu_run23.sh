#!/bin/bash
# sanitizer passes over the round-2 kernel exercise (memcheck, racecheck with the pair kernel and with MN_TC_CG=1)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_round2_kernels.py > gpurun_out/r2_sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_round2_kernels.py > gpurun_out/r2_sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.txt
MN_TC_CG=1 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_round2_kernels.py > gpurun_out/r2_sanitizer_racecheck_cg1.txt 2>&1
echo "racecheck cg1 rc=$?" >> gpurun_out/r2_sanitizer_racecheck_cg1.txt
tail -n 5 gpurun_out/r2_sanitizer_memcheck.txt; tail -n 5 gpurun_out/r2_sanitizer_racecheck.txt; tail -n 5 gpurun_out/r2_sanitizer_racecheck_cg1.txt
