#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run14.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -q -x --timeout 300 -p no:cacheprovider >> $O 2>&1
rc=$?
echo "tests rc=$rc" >> $O
if [ $rc -eq 0 ]; then
  for cfg in "MN_CONV_NARROW=1" "MN_CONV_NARROW=0"; do
    echo "== $cfg" >> $O
    env $cfg timeout 300 python tools/profile_sections.py >> $O 2>&1
  done
fi
tail -n 12 $O
