python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py --no-extra > gpurun_out/k_f2.json 2>/dev/null
CASSIE_B200_SYNCMASK=42 python bench.py --no-extra > gpurun_out/k_f2_sync42.json 2>/dev/null
CASSIE_B200_SYNCMASK=40 python bench.py --no-extra > gpurun_out/k_f2_sync40.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k_f2*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); e=d['e2e']
    print(f, round(d['value']/1e6,2), round(d['multi_tick_launches']['env_steps_per_s']/1e6,2), round(e['value']/1e6,2), d['roofline']['issue']['frac'])
PY
