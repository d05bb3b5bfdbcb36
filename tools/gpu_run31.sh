#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider --timeout 600 > gpurun_out/r2_pytest_gpu_fusegn.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_fusegn.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_fusegn.txt 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_smoke_fusegn.txt
grep -c PASSED gpurun_out/r2_pytest_gpu_fusegn.txt; tail -n 4 gpurun_out/r2_pytest_gpu_fusegn.txt; tail -n 3 gpurun_out/r2_smoke_fusegn.txt
