#!/bin/bash
# Round-2 evidence on 8 B200 of one node (run under gpurun --gpus 8 from the repo root): the scaling lines the driver also takes, for configs 2 and 3
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus 8 --no-extra > gpurun_out/r2_bench_8gpu_config2.json 2> gpurun_out/r2_bench_8gpu_config2.err
$TR bench.py --gpus 8 --config 3 > gpurun_out/r2_bench_8gpu_config3.json 2> gpurun_out/r2_bench_8gpu_config3.err
if [ -n "$WITH_REFERENCE_ARM" ]; then $TR bench.py --gpus 8 --impl reference > gpurun_out/r2_bench_reference_arm_8gpu_box.json 2> gpurun_out/r2_bench_reference_arm_8gpu_box.err; fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_8gpu_config*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); e=d['e2e']
    print(f, d['n_gpus'], round(d['value']/1e6,2), round(e['value']/1e6,2), e.get('host_threads'), d['per_rank']['e2e_ms_per_step'], d['clocks'])
PY
