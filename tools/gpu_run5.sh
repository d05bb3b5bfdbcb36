#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run5.txt
: > $O
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x --timeout 300 -p no:cacheprovider >> $O 2>&1
rc=$?
echo "tc tests rc=$rc" >> $O
if [ $rc -eq 0 ]; then
  for shape in "16 128 128 128 128 3" "1 128 2048 128 64 3" "1 128 2048 64 64 3" "1 8 512 256 256 3" "16 32 32 512 512 3"; do
    timeout 120 python tools/bench_conv.py $shape >> $O 2>&1
  done
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_tc2_c128 python tools/bench_conv.py 16 128 128 128 128 3 1 3 >> $O 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_tc2_c512 python tools/bench_conv.py 16 32 32 512 512 3 1 3 >> $O 2>&1
  timeout 300 python tools/profile_sections.py >> $O 2>&1
fi
cat $O | tail -n 30
