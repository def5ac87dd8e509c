#!/bin/bash
# Round-2 evidence on one B200 (run under gpurun from the repo root): GPU tests, ncu captures of the step kernel, the static counts bench.py quotes,
# the ncu launch list of bench.py, then the bench lines of every BASELINE configuration and the reference arm.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_gpu_tests.txt; cat gpurun_out/r2_gpu_tests.txt
NCU="ncu --set full --clock-control none --import-source on -k regex:cassie_step_kernel -s 8 -c 1"
$NCU -o gpurun_out/prof_r2z_plain python tools/dev/prof_step.py 2 > gpurun_out/prof_r2z_plain.log 2>&1
EST=1 $NCU -o gpurun_out/prof_r2z_est python tools/dev/prof_step.py 2 > gpurun_out/prof_r2z_est.log 2>&1
$NCU -o gpurun_out/prof_r2z_cfg3 python tools/dev/prof_step.py 3 > gpurun_out/prof_r2z_cfg3.log 2>&1
python tools/ncu_counts.py gpurun_out/prof_r2z_plain.ncu-rep 4096 2 > gpurun_out/counts_a.json && python tools/ncu_counts.py gpurun_out/prof_r2z_cfg3.ncu-rep 16384 3 gpurun_out/counts_a.json > gpurun_out/r2_step_kernel_counts.json \
  && cp gpurun_out/r2_step_kernel_counts.json profiles/r2_step_kernel_counts.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/r2_launches_bench.log 2>&1
python bench.py > gpurun_out/r2_bench_1gpu_config2.json 2> gpurun_out/r2_bench_1gpu_config2.err
for c in 3 4 5; do python bench.py --config $c > gpurun_out/r2_bench_1gpu_config$c.json 2> gpurun_out/r2_bench_1gpu_config$c.err; done
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm_1gpu_box.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_1gpu_config*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); e=d['e2e']
    print(f, round(d['value']/1e6,2), round(d['multi_tick_launches']['env_steps_per_s']/1e6,2), round(e['value']/1e6,2), (d['roofline'].get('issue') or {}).get('frac'), d['clocks'])
PY
