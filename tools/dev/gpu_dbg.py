"""dev aid: a short tour of every kernel instance / mode, meant to be run under compute-sanitizer (not a pytest file)"""
import importlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from conftest import PD_TARGET, PD_PGAIN, PD_DGAIN
from test_task_pd import task_rows
P = importlib.import_module('cassie-mujoco-sim_b200')
for prec in (P.FP64, P.FP32):
    for model in ('cassie', 'cassie_tray_box', 'cassie_hfield'):
        b = P.CassieBatch(37, modelfile=P.model_path(model), precision=prec)
        b.set_pd(P.pd_rows(37, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
        b.step(1); b.step(7); b.forward(); b.integrate_pos()
        b.enable_aux(); b.step(1); b.step(5); b.query(); a = b.aux()
        b.set_model('dof_damping', b.get_model('dof_damping') * 1.1); b.set_const(reset_state=True); b.step(3)
        b.set_task_pd(np.tile(task_rows(np.random.default_rng(1)), (37, 1))); b.step(1); b.step(4)
        m = np.zeros(37, dtype=np.uint8); m[5] = 1; b.reset(m); b.step(2)
        pin = (P.pd_in_t * 37)(); ys = b.step_pd(pin)
        print(prec, model, 'ok', np.isfinite(b.qpos()).all(), a[0][2])
