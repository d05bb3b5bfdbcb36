#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run8.txt
: > $O
echo "== bench default" >> $O
timeout 600 python bench.py --steps 10 --warmup 3 --no-collective --no-cpu-baseline > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
echo "== bench MN_FUSE_GN=1" >> $O
MN_FUSE_GN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-collective --no-cpu-baseline > gpurun_out/r2_bench_e_fusegn.json 2> gpurun_out/r2_bench_e_fusegn.err
python - >> $O <<'PY'
import json
for f in ('gpurun_out/r2_bench_e.json','gpurun_out/r2_bench_e_fusegn.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],3), 'module ms', round(d['config']['eager_ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
MN_FUSE_GN=1 timeout 300 python tools/profile_sections.py >> $O 2>&1
timeout 300 python tools/profile_sections.py >> $O 2>&1
# launch list of exactly one step (module API), cold-cache serialised: SHARES only
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --profile --steps 1 --warmup 3 --no-collective > gpurun_out/r2_launches_step.out 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2_launches_step.csv)" >> $O
# full capture of the roofline kernel at 16 chars (current build)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_tc2_c512_final python tools/bench_conv.py 16 32 32 512 512 3 1 3 > /dev/null 2>&1
echo "ncu full rc=$?" >> $O
cat $O | tail -n 30
