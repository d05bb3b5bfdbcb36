#!/bin/bash
# Round-2 final evidence run (single B200): full GPU test suite incl. the staged unmodified reference scripts, smoke, bench with the
# driver's arguments, reference arm, configs[4] sweep, section / per-layer timings, ncu launch lists (cold and warm) + full capture of
# the roofline kernel.  (compute-sanitizer runs separately: tools/gpu_run23.sh.)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider --timeout 600 > gpurun_out/r2_pytest_gpu_final.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_final.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_smoke.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
echo "reference arm rc=$?"
timeout 300 python tools/bench_style_sweep.py --mode batched > gpurun_out/r2_style_sweep_batched.json 2> gpurun_out/r2_style_sweep.err
timeout 300 python tools/bench_style_sweep.py --mode per_step > gpurun_out/r2_style_sweep_per_step.json 2>> gpurun_out/r2_style_sweep.err
timeout 300 python tools/profile_sections.py > gpurun_out/r2_sections_final.json 2>&1
MN_MODULE_GRAPHS=0 timeout 300 python tools/profile_conv_layers.py > gpurun_out/r2_conv_layers.txt 2>&1
timeout 300 python tools/bench_small_convs.py > gpurun_out/r2_small_convs_final.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --profile --steps 1 --warmup 3 --no-collective > gpurun_out/r2_launches_step.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_warm.csv python bench.py --profile --steps 1 --warmup 3 --no-collective > gpurun_out/r2_launches_step_warm.out 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_tc2_c512_final python tools/bench_conv.py 16 32 32 512 512 3 1 3 > /dev/null 2>&1
tail -n 4 gpurun_out/r2_pytest_gpu_final.txt; tail -n 3 gpurun_out/r2_smoke.txt; cat gpurun_out/r2_sections_final.json | tail -n 1 | cut -c1-600
