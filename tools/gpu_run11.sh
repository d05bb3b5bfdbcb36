#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run11.txt
: > $O
for cfg in "MN_TC_HALO_STAGES=0" "MN_TC_HALO_STAGES=3"; do
  echo "== $cfg" >> $O
  for shape in "1 128 2048 64 64 3" "1 128 2048 128 64 3" "16 128 128 128 128 3" "1 8 512 256 256 3" "1 8 512 128 128 3" "1 16 512 64 64 3"; do
    env $cfg timeout 120 python tools/bench_conv.py $shape >> $O 2>&1
  done
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_tc2_c64 python tools/bench_conv.py 1 128 2048 64 64 3 1 3 > /dev/null 2>&1
echo "ncu rc=$?" >> $O
cat $O
