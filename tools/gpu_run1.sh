#!/bin/bash
# Round-2 GPU check #1: full GPU test suite (incl. the staged unmodified reference scripts), MMA probes, stock-PyTorch baseline,
# style sweep, bench.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
./tools/mma_probe2 > gpurun_out/r2_mma_probe2.txt 2>&1
./tools/mma_probe > gpurun_out/r2_mma_probe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -rs > gpurun_out/r2_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.txt
timeout 300 python tests/gpu_torch_baseline.py --compare > gpurun_out/r2_torch_cuda_baseline.json 2> gpurun_out/r2_torch_cuda_baseline.err
timeout 300 python tools/bench_style_sweep.py --mode batched > gpurun_out/r2_style_sweep_batched.json 2> gpurun_out/r2_style_sweep.err
timeout 300 python tools/bench_style_sweep.py --mode per_step > gpurun_out/r2_style_sweep_per_step.json 2>> gpurun_out/r2_style_sweep.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -n 30 gpurun_out/r2_pytest_gpu.txt
cat gpurun_out/r2_mma_probe2.txt
