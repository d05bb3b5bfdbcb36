#!/bin/bash
# 2 x B200: the two-GPU tests (peer exchange, wide parity) and the driver-style torchrun bench of the final build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer_exchange.py tests/test_gpu_parity_wide.py -m gpu -q -rA --timeout 600 -p no:cacheprovider > gpurun_out/r2_pytest_2gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 3 > gpurun_out/r2_bench_2gpu_reference_arm.json 2> gpurun_out/r2_bench_2gpu_reference_arm.err
echo "reference arm rc=$?"
tail -n 6 gpurun_out/r2_pytest_2gpu.txt; tail -c 1500 gpurun_out/r2_bench_2gpu.json; tail -c 400 gpurun_out/r2_bench_2gpu_reference_arm.json
