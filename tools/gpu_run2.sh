#!/bin/bash
# Round-2 GPU check #2: cta_group::2 conv kernel -- numerics first, then A/B timing against the cta_group::1 path.
mkdir -p gpurun_out
O=gpurun_out/r2_run2.txt
: > $O
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -q -x --timeout 300 -p no:cacheprovider >> $O 2>&1
rc=$?
echo "tc tests rc=$rc" >> $O
if [ $rc -eq 0 ]; then
  for cg in 0 1; do
    echo "== MN_TC_CG=$cg" >> $O
    for shape in "16 32 32 512 512 3" "128 32 32 512 512 3" "1 64 1024 256 256 3" "16 64 64 256 256 3" "16 128 128 128 128 3" "1 128 2048 64 64 3" "1 8 512 512 512 3" "1 8 512 256 256 3" "16 32 32 256 256 3"; do
      MN_TC_CG=$cg timeout 120 python tools/bench_conv.py $shape >> $O 2>&1
    done
    MN_TC_CG=$cg timeout 300 python tools/profile_sections.py >> $O 2>&1
  done
  timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r2_run2_pytest.txt 2>&1
  echo "full pytest rc=$?" >> $O
  tail -n 3 gpurun_out/r2_run2_pytest.txt >> $O
  timeout 600 python bench.py --steps 10 --warmup 3 --no-collective > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
  MN_TC_CG=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-collective --no-cpu-baseline > gpurun_out/r2_bench_b_cg1.json 2> gpurun_out/r2_bench_b_cg1.err
fi
cat $O | tail -n 60
