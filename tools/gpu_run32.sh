#!/bin/bash
# same-box A/B at the driver's arguments: previous library (one lane per halo row, GroupNorm apply passes) vs current (four lanes per
# row, fused GroupNorm), alternating twice
mkdir -p gpurun_out
O=gpurun_out/r2_run32.txt
: > $O
cp marconet_b200/libmarconet_b200.so /tmp/libnew.so
run() {  # $1 label, $2 MN_FUSE_GN
  MN_FUSE_GN=$2 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-collective --no-cpu-baseline > gpurun_out/r2_ab_$1.json 2> gpurun_out/r2_ab_$1.err
  python - $1 >> $O <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2_ab_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'module ms', round(d['config']['eager_ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'roof', round(d['roofline']['tensor_pipe_frac'],3), round(d['roofline']['same_kernel_128_chars']['tensor_pipe_frac'],3), d['clocks']['sm_mhz'])
PY
}
for rep in 1 2; do
  cp _ab/libold.so marconet_b200/libmarconet_b200.so; run old_$rep 0
  cp /tmp/libnew.so marconet_b200/libmarconet_b200.so; run new_fused_$rep 1
  run new_unfused_$rep 0
done
cat $O
