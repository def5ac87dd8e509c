#!/bin/bash
# multi-rank check of the end-of-round build on 2 B200 of one node (run under gpurun --gpus 2): the scaling line the driver also takes, config 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus 2 --no-extra > gpurun_out/r2_bench_2gpu_config2.json 2> gpurun_out/r2_bench_2gpu_config2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_config2.json').read().strip().splitlines()[-1]); e=d['e2e']
print(d['n_gpus'], round(d['value']/1e6,2), round(e['value']/1e6,2), e.get('host_threads'), e.get('split_this_rank_ms'), d['clocks'])
PY
tail -3 gpurun_out/r2_bench_2gpu_config2.err
