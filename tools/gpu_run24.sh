#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_warm.csv python bench.py --profile --steps 1 --warmup 3 --no-collective > gpurun_out/r2_launches_step_warm.out 2>&1
echo rc=$?
wc -l gpurun_out/r2_launches_step_warm.csv
