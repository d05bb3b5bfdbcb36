#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run22.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_tc.py -q -x --timeout 300 -p no:cacheprovider >> $O 2>&1
rc=$?
echo "tests rc=$rc" >> $O
timeout 300 python tools/bench_small_convs.py >> $O 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-collective --no-cpu-baseline > gpurun_out/r2_bench_l.json 2> gpurun_out/r2_bench_l.err
python - >> $O <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_l.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],3), 'module ms', round(d['config']['eager_ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['tensor_pipe_frac'],3))
PY
timeout 300 python tools/profile_sections.py >> $O 2>&1
tail -n 26 $O | cut -c1-330
