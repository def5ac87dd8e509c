python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py --no-extra > gpurun_out/k_kids.json 2>/dev/null
CASSIE_B200_NOKIDS=1 python bench.py --no-extra > gpurun_out/k_nokids.json 2>/dev/null
CASSIE_B200_AOS_THREADS=12 python bench.py --no-extra > gpurun_out/k_kids_t12.json 2>/dev/null
CASSIE_B200_AOS_THREADS=8 python bench.py --no-extra > gpurun_out/k_kids_t8.json 2>/dev/null
python bench.py --config 3 > gpurun_out/k3_kids.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k*kids*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); e=d['e2e']
    print(f, round(d['value']/1e6,2), round(d['multi_tick_launches']['env_steps_per_s']/1e6,2), round(e['value']/1e6,2), e['split_this_rank_ms'])
PY
