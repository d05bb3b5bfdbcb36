#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run6.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_models.py -q -x --timeout 300 -p no:cacheprovider >> $O 2>&1
rc=$?
echo "tests rc=$rc" >> $O
if [ $rc -eq 0 ]; then
  for shape in "16 32 32 512 512 3" "1 64 1024 256 256 3" "16 128 128 256 128 3" "16 128 128 128 128 3" "1 128 2048 128 64 3" "1 128 2048 64 64 3" "16 32 32 256 256 3" "1 8 512 256 256 3" "1 8 512 512 512 3"; do
    timeout 120 python tools/bench_conv.py $shape >> $O 2>&1
  done
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_tc2_c128_b python tools/bench_conv.py 16 128 128 128 128 3 1 3 > /dev/null 2>&1
  timeout 300 python tools/profile_sections.py >> $O 2>&1
  timeout 600 python bench.py --steps 10 --warmup 3 --no-collective > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
fi
cat $O | tail -n 30
