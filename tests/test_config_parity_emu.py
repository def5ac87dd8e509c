"""CPU rehearsal of tests/test_gpu_configs.py: the product's stepper source in the host emulation (fp32 and fp64 instances) against the oracle on a few
environments of each BASELINE configuration, with bench.py's own seeds, push schedule, terrains and gaits.  fp64 <= 1e-10, fp32 <= 1e-4 over 400 ticks."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO

sys.path.insert(0, REPO)
import bench  # noqa: E402

PD_ROW = np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN])


@pytest.mark.parametrize('cfg', [3, 4, 5])
def test_emulated_kernel_follows_the_oracle_on_config(oracle_mod, cfg):
    import emu_harness as EH
    O = oracle_mod
    model, n = bench.CONFIGS[cfg]['model'], bench.CONFIGS[cfg]['envs']
    om, cm = os.path.join(GOLDEN, model + '.omodel'), os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', model + '.cmodel')
    terr = None
    if cfg == 4:
        terr = (np.random.default_rng(7).uniform(0, 1, (64, 200, 200)) * 0.25).astype(np.float32)
        terr[:, 95:105, 95:105] = 0
    o0 = O.OracleSim(om)
    nq = o0.nq
    Q = bench.jittered_qpos(o0.arr('qpos')[:nq].copy(), n, seed=0)
    f, ph = bench.philox_uniform(99, 0, n, 0.5, 1.5), bench.philox_uniform(99, 1, n, 0.0, 2 * np.pi)
    for e, fp32 in ((7, False), (n - 3, True)):
        o, em = O.OracleSim(om), EH.EmuSim(cm, fp32=fp32)
        em.plain()
        if terr is not None:
            t = np.ascontiguousarray(terr[e % 64])
            C.memmove(o.L.osim_hfield_data(o.h), t.ctypes.data, t.nbytes)
            em.set_hfield(t)
        o.arr('qpos')[:nq] = Q[e]
        o.forward()
        qq = em.get('qpos')
        qq[:nq] = Q[e]
        em.set('qpos', qq)
        em.forward()
        phase = ph[e] + np.array([0.0] * 5 + [np.pi] * 5)
        if cfg == 5:
            em.set_gait(np.array(bench.GAIT_AMP), phase, f[e])
        worst = 0.0
        for t in range(400):
            if cfg == 3:
                if t % 400 == 0:
                    push = bench.philox_uniform(1234, t // 400, (n, 2), -100.0, 100.0)[e]
                    o.arr('xfrc_applied').reshape(-1, 6)[1, :2] = push
                    x = np.zeros(8); x[:2] = push; x[6] = 1
                    em.set('xfrc', x)
                elif t % 400 == 100:
                    o.arr('xfrc_applied')[:] = 0
                    em.set('xfrc', np.zeros(8))
            tgt = np.array(PD_TARGET) + (np.array(bench.GAIT_AMP) * np.sin(2 * np.pi * f[e] * t * 0.0005 + phase) if cfg == 5 else 0)
            o.step_pd(O.make_pd(pTarget=tgt, pGain=PD_PGAIN, dGain=PD_DGAIN))
            em.step(PD_ROW, 1)
            worst = max(worst, float(np.abs(em.get('qpos')[:nq] - o.arr('qpos')[:nq]).max()))
        assert worst < (1e-4 if fp32 else 1e-10), (cfg, e, fp32, worst)
