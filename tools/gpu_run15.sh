#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run15.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_graph.py -q -x --timeout 300 -p no:cacheprovider >> $O 2>&1
rc=$?
echo "tests rc=$rc" >> $O
if [ $rc -eq 0 ]; then
  for cfg in "MN_TC_TAIL=1" "MN_TC_TAIL=0"; do
    echo "== $cfg" >> $O
    for shape in "16 32 32 512 512 3" "16 32 32 256 256 3" "1 64 1024 256 128 3" "16 16 16 512 512 3" "16 64 64 256 256 3"; do
      env $cfg timeout 120 python tools/bench_conv.py $shape >> $O 2>&1
    done
    env $cfg timeout 600 python bench.py --steps 10 --warmup 3 --no-collective --no-cpu-baseline > gpurun_out/r2_bench_i_$cfg.json 2> gpurun_out/r2_bench_i_$cfg.err
    python - "$cfg" >> $O <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2_bench_i_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],3), 'module ms', round(d['config']['eager_ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'roof', round(d['roofline']['tensor_pipe_frac'],3), round(d['roofline']['same_kernel_128_chars']['tensor_pipe_frac'],3))
PY
  done
fi
tail -n 22 $O
