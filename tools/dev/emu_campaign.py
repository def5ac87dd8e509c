"""dev aid: long randomised emu-vs-oracle campaign (CPU only).  usage: emu_campaign.py [seeds] [ticks]"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (REPO, os.path.join(REPO, 'oracle'), os.path.join(REPO, 'tests')):
    sys.path.insert(0, p)
import oracle as O
import emu_harness as E
from conftest import PD_TARGET, PD_PGAIN, PD_DGAIN, GOLDEN
nseeds, ticks = (int(sys.argv[1]) if len(sys.argv) > 1 else 5), (int(sys.argv[2]) if len(sys.argv) > 2 else 3000)
worst_all = 0.0
for model in ('cassie', 'cassie_tray_box', 'cassie_hfield'):
    for seed in range(nseeds):
        rng = np.random.default_rng(1000 + seed)
        o = O.OracleSim(os.path.join(GOLDEN, model + '.omodel'))
        e = E.EmuSim(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', model + '.cmodel'))
        if seed % 2:
            e.plain()
        if model == 'cassie_hfield':
            h = (rng.random((200, 200)) * rng.uniform(0.05, 0.5)).astype(np.float32)
            np.ctypeslib.as_array(o.L.osim_hfield_data(o.h), shape=(40000,))[:] = h.ravel(); e.set_hfield(h)
        pelvis = 2 if model == 'cassie_hfield' else 1
        worst, rows, cross, drops, mism = 0.0, set(), 0, 0, 0
        for k in range(ticks):
            if k % 80 == 0:
                tgt = np.array(PD_TARGET) + rng.uniform(-0.4, 0.4, 10)
                if rng.random() < 0.3: tgt[0], tgt[5] = -0.26, 0.26
                g = rng.uniform(0.3, 1.5)
                u = O.make_pd(pTarget=tgt, pGain=np.array(PD_PGAIN) * g, dGain=PD_DGAIN)
                pd = np.concatenate([np.zeros(10), tgt, np.zeros(10), np.array(PD_PGAIN) * g, PD_DGAIN])
            if k % 200 == 0:
                f = np.zeros(6); f[:3] = rng.uniform(-250, 250, 3); f[3:] = rng.uniform(-30, 30, 3)
                o.arr('xfrc_applied').reshape(-1, 6)[pelvis] = f; e.set('xfrc', np.concatenate([f, [pelvis, 0]]))
            if k % 200 == 50:
                o.arr('xfrc_applied')[:] = 0; e.set('xfrc', np.zeros(8))
            o.step_pd(u); e.step(pd)
            c = e.get('counters'); rows.add(int(c[0])); cross += o.check_self_collision(); drops += int(c[4]) > 0
            if k % 50 == 0:
                d = np.abs(e.get('qpos')[:o.nq] - o.arr('qpos')).max()
                if int(c[4]) == 0 and o.get_int('dropped_contacts') == 0: worst = max(worst, d)
                else: mism += 1
        worst_all = max(worst_all, worst if drops == 0 else 0)
        print('%-16s seed %d  max|dqpos| %.2e  rows %d..%d  leg-leg ticks %d  ticks with dropped contacts %d  z %.2f' % (model, seed, worst, min(rows), max(rows), cross, drops, o.arr('qpos')[2]), flush=True)
print('worst over runs without dropped contacts: %.2e' % worst_all)
