#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run7.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_module_graphs.py -q -x --timeout 300 -p no:cacheprovider >> $O 2>&1
echo "module graph tests rc=$?" >> $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r2_run7_pytest.txt 2>&1
echo "full pytest rc=$?" >> $O
tail -n 15 gpurun_out/r2_run7_pytest.txt >> $O
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
echo "bench rc=$?" >> $O
tail -n 5 gpurun_out/r2_bench_d.err >> $O
cat $O | tail -n 40
