L=cassie-mujoco-sim_b200/libcassie_b200.so; A=cassie-mujoco-sim_b200/libcassie_b200_alt.so
cp $L /tmp/main.so
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do
  cp /tmp/main.so $L; python bench.py --no-extra > gpurun_out/dbg_new_$i.json 2>/dev/null
  cp $A $L; python bench.py --no-extra > gpurun_out/dbg_old_$i.json 2>/dev/null
done
cp /tmp/main.so $L
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/dbg_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); e=d['e2e']
    print(f, round(d['value']/1e6,2), round(d['multi_tick_launches']['env_steps_per_s']/1e6,2), round(e['value']/1e6,2), round(e['estimator_off_variant']['value']/1e6,2))
PY
