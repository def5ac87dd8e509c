python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --no-extra > gpurun_out/k_stage.json 2>/dev/null
python bench.py --config 3 > gpurun_out/k3_stage.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k*_stage.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); e=d['e2e']
    print(f, round(d['value']/1e6,2), d.get('multi_tick_launches',{}).get('value'), round(e['value']/1e6,2), e.get('split_this_rank_ms'))
PY
