#!/bin/bash
# last check of the round on the committed tree: the driver's own commands
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA -p no:cacheprovider --timeout 600 > gpurun_out/r2_pytest_gpu_final.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_smoke.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
echo "bench rc=$?"
tail -n 3 gpurun_out/r2_pytest_gpu_final.txt; tail -n 2 gpurun_out/r2_smoke.txt; tail -c 300 gpurun_out/r2_bench_final.json
