#!/bin/bash
# BASELINE config 4 at its stated size (4 GPUs x 8192 envs on the height field), run under gpurun --gpus 4 from the repo root
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --config 4 > gpurun_out/r2_bench_4gpu_config4.json 2> gpurun_out/r2_bench_4gpu_config4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_4gpu_config4.json').read().strip().splitlines()[-1]); e=d['e2e']
print(d['n_gpus'], round(d['value']/1e6,2), round(e['value']/1e6,2), e.get('host_threads'), d['per_rank']['e2e_ms_per_step'], d['clocks'])
PY
