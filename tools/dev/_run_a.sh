python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for d in 0 1; do for c in 2 4; do CASSIE_B200_AOS_OBS_DMA=$d CASSIE_B200_AOS_CHUNKS=$c python bench.py --no-extra > gpurun_out/e2e_c2_dma${d}_chunks$c.json 2>/dev/null; done; done
for d in 0 1; do CASSIE_B200_AOS_OBS_DMA=$d python bench.py --config 3 > gpurun_out/e2e_c3_dma${d}.json 2>/dev/null; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e2e_*dma*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); e=d['e2e']
    print(f, round(d['value']/1e6,2), d.get('multi_tick_launches',{}).get('value'), round(e['value']/1e6,2), e.get('split_this_rank_ms'))
PY
