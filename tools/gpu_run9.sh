#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run9.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_precision_plan.py tests/test_gpu_models.py tests/test_gpu_graph.py tests/test_gpu_module_graphs.py -q -x --timeout 300 -p no:cacheprovider >> $O 2>&1
rc=$?
echo "tests rc=$rc" >> $O
if [ $rc -eq 0 ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --no-collective --no-cpu-baseline > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
  python - >> $O <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_f.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],3), 'module ms', round(d['config']['eager_ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
PY
  timeout 300 python tools/profile_sections.py >> $O 2>&1
fi
tail -n 25 $O
