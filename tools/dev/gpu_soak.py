"""GPU soak (development aid, not a pytest file): random PD targets re-drawn every 50 ticks, pushes, 16384 envs, 4000 ticks; everything must stay finite."""
import importlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from conftest import PD_TARGET, PD_PGAIN, PD_DGAIN
P = importlib.import_module('cassie-mujoco-sim_b200')
rng = np.random.default_rng(5)
for name, n in (('cassie', 16384), ('cassie_hfield', 4096), ('cassie_tray_box', 4096)):
    b = P.CassieBatch(n, modelfile=P.model_path(name), precision=P.FP32)
    if name == 'cassie_hfield':
        T = (rng.uniform(0, 1, (16, 200, 200)) * 0.5).astype(np.float32); b.set_hfield_data(T)
    worst_rows = 0; dropped = 0
    for it in range(80):
        tgt = np.array(PD_TARGET) + rng.uniform(-0.4, 0.4, (n, 10)) * np.array([0.3, 0.3, 1, 1, 1] * 2)
        b.set_pd(P.pd_rows(n, pTarget=tgt, pGain=PD_PGAIN, dGain=PD_DGAIN))
        if it % 8 == 3:
            f = np.zeros((n, 6)); f[:, :2] = rng.uniform(-150, 150, (n, 2)); b.apply_force(f, 'cassie-pelvis')
        if it % 8 == 5:
            b.clear_forces()
        b.step(50)
        if it % 20 == 19:      # reset the fallen ones (pelvis below 0.4 m), like an RL loop would
            q = b.qpos(); c = b.counters()
            assert np.isfinite(q).all() and np.isfinite(b.qvel()).all(), (name, it)
            worst_rows = max(worst_rows, int(c[:, 0].max())); dropped += int(c[:, 4].sum())
            b.reset((q[:, 2] < 0.4).astype(np.uint8))
    q = b.qpos()
    print(name, 'ok: finite after 4000 ticks; max rows', worst_rows, 'dropped contacts (cumulative)', dropped, 'fallen at end', int((q[:, 2] < 0.4).sum()), 'of', n, flush=True)
    b.close()
