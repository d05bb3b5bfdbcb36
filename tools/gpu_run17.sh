#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run17.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_precision_plan.py -q -x --timeout 300 -p no:cacheprovider >> $O 2>&1
rc=$?
echo "tests rc=$rc" >> $O
if [ $rc -eq 0 ]; then
  for shape in "16 32 32 512 512 3" "1 64 1024 256 256 3" "16 128 128 128 128 3" "1 128 2048 128 64 3" "1 128 2048 64 64 3" "1 8 512 256 256 3" "1 8 512 512 512 3" "1 16 512 64 64 3"; do
    timeout 120 python tools/bench_conv.py $shape >> $O 2>&1
  done
  timeout 120 python tools/trace_tc2.py 1 128 2048 64 64 > gpurun_out/r2_trace_c64_c.txt 2>&1
  timeout 120 python tools/trace_tc2.py 16 128 128 128 128 > gpurun_out/r2_trace_c128_c.txt 2>&1
  timeout 600 python bench.py --steps 10 --warmup 3 --no-collective --no-cpu-baseline > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err
  python - >> $O <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_k.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],3), 'module ms', round(d['config']['eager_ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['tensor_pipe_frac'],3), round(d['roofline']['same_kernel_128_chars']['tensor_pipe_frac'],3))
PY
  timeout 300 python tools/profile_sections.py >> $O 2>&1
fi
tail -n 16 $O; tail -n 3 gpurun_out/r2_trace_c64_c.txt | cut -c1-250
