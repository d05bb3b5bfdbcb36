#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run28.txt
: > $O
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider --deselect tests/test_dropin_scripts.py >> $O 2>&1
echo "tests rc=$?" >> $O
timeout 600 python bench.py --steps 10 --warmup 3 --no-collective --no-cpu-baseline > gpurun_out/r2_bench_m.json 2> gpurun_out/r2_bench_m.err
python - >> $O <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_m.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],3), 'module ms', round(d['config']['eager_ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['tensor_pipe_frac'],3))
PY
timeout 300 python tools/profile_sections.py >> $O 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_warm2.csv python bench.py --profile --steps 1 --warmup 3 --no-collective > /dev/null 2>&1
tail -n 8 $O | cut -c1-700
