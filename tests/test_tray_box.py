"""BASELINE config 5 ingredient: cassie_tray_box.xml -- an extra free body (the cup, nq 42 / nv 38) riding on a tray welded to the pelvis,
box contacts by our own analytic rules.  Product vs oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO, product

OMODEL = os.path.join(GOLDEN, 'cassie_tray_box.omodel')
CMODEL = os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie_tray_box.cmodel')
PD_ROW = np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN])


def test_emulated_kernel_matches_oracle(oracle_mod):
    import emu_harness as E
    o, e = oracle_mod.OracleSim(OMODEL), E.EmuSim(CMODEL)
    assert o.get_int('nq') == 42 and o.get_int('nv') == 38
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    rode = 0
    for k in range(1500):
        o.step_pd(u)
        e.step(PD_ROW)
        q = o.arr('qpos')
        rode += abs((q[37] - q[2]) - 0.225) < 2e-3          # cup resting on the tray: 0.17 + 0.005 + 0.05 above the pelvis origin
        if k % 50 == 0 or k == 1499:
            assert np.abs(e.get('qpos')[:42] - q).max() < 1e-9, k
            assert int(e.get('counters')[0]) == o.get_int('nefc')
    assert rode > 500 and o.get_int('unsupported_pairs') == 0


def test_cup_free_fall_and_floor_contact(oracle_mod):
    """the cup alone: push it off the tray; it falls freely (z'' = -g) and comes to rest on the floor plane (z = -0.01) on 4 corner contacts"""
    o = oracle_mod.OracleSim(OMODEL)
    o.arr('qpos')[35] = 1.0                        # 1 m in front of the robot
    o.forward()
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    z0 = o.arr('qpos')[37]
    for _ in range(200):
        o.step_pd(u)
    t = 0.1
    assert abs(o.arr('qpos')[37] - (z0 - 0.5 * 9.81 * t * t)) < 5e-4 * 9.81 * t      # semi-implicit Euler: O(h) position error
    for _ in range(1800):                           # (the robot itself topples forward onto the cup a little later)
        o.step_pd(u)
    assert abs(o.arr('qpos')[37] - (-0.01 + 0.05)) < 2e-3 and np.abs(o.arr('qvel')[32:38]).max() < 1e-2


@pytest.mark.gpu
def test_gpu_matches_oracle(oracle_mod):
    P = product()
    n = 4
    b = P.CassieBatch(n, modelfile=CMODEL, precision=P.FP64)
    assert b.nq == 42 and b.nv == 38
    b.set_pd(P.pd_rows(n, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
    o = oracle_mod.OracleSim(OMODEL)
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    for k in range(1500):
        o.step_pd(u)
        b.step(1)
        if k % 100 == 99:
            assert np.abs(b.qpos()[n - 1] - o.arr('qpos')).max() < 1e-9, k
            assert np.abs(b.qvel()[0] - o.arr('qvel')).max() < 1e-7, k
    f = P.CassieBatch(256, modelfile=CMODEL, precision=P.FP32)
    f.set_pd(P.pd_rows(256, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
    f.step(600)
    q = f.qpos()
    assert np.isfinite(q).all() and np.abs(q[:, 37] - q[:, 2] - 0.225).max() < 3e-3      # every cup still rides its tray
    # integrate-pos kernel with the free joint
    v = np.random.default_rng(0).normal(size=(256, 38))
    q0 = f.qpos(); f.set_qvel(v); f.integrate_pos(); f.sync(); q1 = f.qpos()
    assert np.abs(q1[:, 35:38] - (q0[:, 35:38] + 5e-4 * v[:, 32:35])).max() < 1e-6
    assert np.abs(np.linalg.norm(q1[:, 38:42], axis=1) - 1).max() < 1e-5
