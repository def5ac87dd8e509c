"""Randomised closed-loop stress of the kernel source (host emulation, fp64) against the oracle: moving PD targets, pelvis pushes, legs driven
together (leg-leg contacts send constraint rows across both legs: the long path of the row transform), falls.  Catches any divergence between
the fast paths (mirrored-leg rows, fused factorisation, dense / reduction solver paths) and the plain algorithm the oracle runs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO


@pytest.mark.parametrize('model,seed', [('cassie', 0), ('cassie', 1), ('cassie_tray_box', 2), ('cassie_hfield', 3)])
def test_random_actions_and_pushes(oracle_mod, model, seed):
    import emu_harness as E
    rng = np.random.default_rng(seed)
    o = oracle_mod.OracleSim(os.path.join(GOLDEN, model + '.omodel'))
    e = E.EmuSim(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', model + '.cmodel'))
    if model == 'cassie_hfield':
        h = (rng.random((200, 200)) * 0.3).astype(np.float32)
        np.ctypeslib.as_array(o.L.osim_hfield_data(o.h), shape=(40000,))[:] = h.ravel()
        e.set_hfield(h)
    nb = o.get_int('nbody')
    pelvis = 1 if model != 'cassie_hfield' else 2
    seen_rows, seen_cross, worst = set(), 0, 0.0
    tgt = np.array(PD_TARGET, dtype=float)
    for k in range(1500):
        if k % 100 == 0:      # new targets; every third segment squeezes the legs together / swings them across
            tgt = np.array(PD_TARGET) + rng.uniform(-0.25, 0.25, 10)
            if (k // 100) % 3 == 2:
                tgt[0], tgt[5] = -0.26, 0.26
                tgt[1], tgt[6] = rng.uniform(-0.35, 0.35), rng.uniform(-0.35, 0.35)
            u = oracle_mod.make_pd(pTarget=tgt, pGain=PD_PGAIN, dGain=PD_DGAIN)
            pd = np.concatenate([np.zeros(10), tgt, np.zeros(10), PD_PGAIN, PD_DGAIN])
        if k % 250 == 0:      # pelvis push for 60 ticks
            f = np.zeros(6)
            f[:3] = rng.uniform(-150, 150, 3)
            o.arr('xfrc_applied').reshape(-1, 6)[pelvis] = f
            e.set('xfrc', np.concatenate([f, [pelvis, 0]]))
        if k % 250 == 60:
            o.arr('xfrc_applied')[:] = 0
            e.set('xfrc', np.zeros(8))
        o.step_pd(u)
        e.step(pd)
        c = e.get('counters')
        seen_rows.add(int(c[0]))
        if k % 25 == 0 or k > 1490:
            worst = max(worst, np.abs(e.get('qpos')[:o.nq] - o.arr('qpos')).max())
            if int(c[4]) == 0:   # no contact was dropped by the product's 12-contact cap: the row sets are the same
                assert int(c[0]) == o.get_int('nefc'), k
        seen_cross += o.check_self_collision()
    assert worst < 1e-7, worst
    assert len(seen_rows) > 3          # the contact state really varied
    assert np.isfinite(o.arr('qpos')).all()


def test_crossed_legs_in_closed_loop(oracle_mod):
    """shins start interpenetrating and are held crossed by the PD targets: rows across both legs for dozens of ticks (until the contact pushes the legs apart), mixed with
    single-leg rows in the same pass"""
    import emu_harness as E
    o = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    e = E.EmuSim(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie.cmodel'))
    q = o.arr('qpos')
    q[7], q[21] = -0.3, 0.3
    qe = e.get('qpos')
    qe[:35] = q
    e.set('qpos', qe)
    o.forward()
    e.forward()
    tgt = np.array(PD_TARGET, dtype=float)
    tgt[0], tgt[5] = -0.26, 0.26
    u = oracle_mod.make_pd(pTarget=tgt, pGain=PD_PGAIN, dGain=PD_DGAIN)
    pd = np.concatenate([np.zeros(10), tgt, np.zeros(10), PD_PGAIN, PD_DGAIN])
    cross, worst = 0, 0.0
    for k in range(600):
        o.step_pd(u)
        e.step(pd)
        cross += o.check_self_collision()
        if k % 20 == 0:
            worst = max(worst, np.abs(e.get('qpos')[:35] - o.arr('qpos')).max())
            assert int(e.get('counters')[0]) == o.get_int('nefc')
    assert cross > 20 and worst < 1e-8, (cross, worst)
