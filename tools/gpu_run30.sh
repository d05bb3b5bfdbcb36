#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run30.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_precision_plan.py -q -x --timeout 300 -p no:cacheprovider >> $O 2>&1
rc=$?
echo "tests rc=$rc" >> $O
if [ $rc -eq 0 ]; then
for shape in "16 32 32 512 512 3" "16 128 128 128 128 3" "1 128 2048 64 64 3"; do
  timeout 120 python tools/bench_conv.py $shape >> $O 2>&1
done
for mode in 0 auto 1; do
  MN_FUSE_GN=$mode timeout 600 python bench.py --steps 10 --warmup 3 --no-collective --no-cpu-baseline > gpurun_out/r2_bench_gn_$mode.json 2> gpurun_out/r2_bench_gn_$mode.err
  python - $mode >> $O <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2_bench_gn_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print('MN_FUSE_GN', sys.argv[1], 'ms', round(d['ms_per_step'],3), 'module ms', round(d['config']['eager_ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'roof', round(d['roofline']['tensor_pipe_frac'],3))
PY
  MN_FUSE_GN=$mode timeout 300 python tools/profile_sections.py >> $O 2>&1
done
fi
tail -n 12 $O | cut -c1-600
