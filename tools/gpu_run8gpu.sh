#!/bin/bash
# 8 x B200: driver-style torchrun bench of the final build (line-sharded value + collective record)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_8gpu.json
