#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run29.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_precision_plan.py -q -x --timeout 300 -p no:cacheprovider -k "gn or GN or groupnorm or fuse" >> $O 2>&1
echo "gn tests rc=$?" >> $O
for mode in 0 auto 1; do
  MN_FUSE_GN=$mode timeout 600 python bench.py --steps 10 --warmup 3 --no-collective --no-cpu-baseline > gpurun_out/r2_bench_gn_$mode.json 2> gpurun_out/r2_bench_gn_$mode.err
  python - $mode >> $O <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2_bench_gn_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print('MN_FUSE_GN', sys.argv[1], 'ms', round(d['ms_per_step'],3), 'module ms', round(d['config']['eager_ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
PY
  MN_FUSE_GN=$mode timeout 300 python tools/profile_sections.py >> $O 2>&1
done
tail -n 8 $O | cut -c1-600
