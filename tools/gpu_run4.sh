#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run4.txt
: > $O
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x --timeout 300 -p no:cacheprovider >> $O 2>&1
rc=$?
echo "tc tests rc=$rc" >> $O
if [ $rc -eq 0 ]; then
  for cfg in "MN_TC_HALO_STAGES=2" "MN_TC_HALO_STAGES=0" "MN_TC_HALO_STAGES=0 MN_TC_EPI_CB=0" "MN_TC_HALO_STAGES=3" ; do
    echo "== $cfg" >> $O
    for shape in "16 32 32 512 512 3" "1 64 1024 256 256 3" "16 128 128 256 128 3" "16 128 128 128 128 3" "1 128 2048 128 64 3" "1 128 2048 64 64 3" "16 32 32 256 256 3" "1 8 512 256 256 3"; do
      env $cfg timeout 120 python tools/bench_conv.py $shape >> $O 2>&1
    done
  done
  timeout 300 python tools/profile_sections.py >> $O 2>&1
fi
cat $O | tail -n 50
