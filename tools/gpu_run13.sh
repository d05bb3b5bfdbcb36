#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run13.txt
: > $O
for cfg in "MN_FUSE_GN=0" "MN_FUSE_GN=1"; do
  echo "== $cfg" >> $O
  env $cfg timeout 600 python bench.py --steps 10 --warmup 3 --no-collective --no-cpu-baseline > gpurun_out/r2_bench_h_$cfg.json 2> gpurun_out/r2_bench_h_$cfg.err
  python - "$cfg" >> $O <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2_bench_h_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],3), 'module ms', round(d['config']['eager_ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
PY
  env $cfg timeout 300 python tools/profile_sections.py >> $O 2>&1
done
cat $O
