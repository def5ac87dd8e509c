python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for c in 2 3 4 5; do python bench.py --config $c --no-extra > gpurun_out/sh_c$c.json 2>/dev/null; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sh_c*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); e=d['e2e']
    print(f, round(d['value']/1e6,2), round(d['multi_tick_launches']['env_steps_per_s']/1e6,2), round(e['value']/1e6,2), round(e['estimator_off_variant']['value']/1e6,2), e['split_this_rank_ms']['device_ms'])
PY
