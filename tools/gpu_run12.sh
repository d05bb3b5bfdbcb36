#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run12.txt
: > $O
for cfg in "MN_TC_NT_ITEMS=120" "MN_TC_NT_ITEMS=60" "MN_TC_NT_ITEMS=30"; do
  echo "== $cfg" >> $O
  for shape in "1 8 512 256 256 3" "1 8 512 128 128 3" "1 8 512 256 256 1" "1 8 512 512 512 1" "1 16 512 64 64 3" "16 16 16 512 512 3" "16 32 32 256 256 3"; do
    env $cfg timeout 120 python tools/bench_conv.py $shape >> $O 2>&1
  done
  env $cfg timeout 300 python tools/profile_sections.py >> $O 2>&1
done
cat $O
