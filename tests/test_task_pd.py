"""pd_in_t's taskPd branch (closed pd_input_step; decoded in the oracle and pinned to the archive by oracle/fuzz_agility.c): the product's
controller stage against the oracle in closed loop."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO

PD_ROW = np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN])


def task_rows(rng):
    """a plausible task-space controller: hold the feet near their standing pose; rows = per leg torque, pTarget, dTarget, pGain, dGain [6]"""
    rows = np.zeros(60)
    for sd in range(2):
        r = rows[30 * sd:30 * sd + 30]
        r[0:6] = rng.uniform(-2, 2, 6)
        r[6:12] = [0.0, 0.135 * (1 - 2 * sd), -0.95, 0, 0, 0] + rng.uniform(-0.02, 0.02, 6)
        r[12:18] = rng.uniform(-0.1, 0.1, 6)
        r[18:24] = [300, 300, 400, 30, 30, 30]
        r[24:30] = [6, 6, 8, 1, 1, 1]
    return rows


def fill_task(u, rows):
    for sd, leg in enumerate((u.leftLeg, u.rightLeg)):
        r = rows[30 * sd:30 * sd + 30]
        for k in range(6):
            leg.taskPd.torque[k], leg.taskPd.pTarget[k], leg.taskPd.dTarget[k], leg.taskPd.pGain[k], leg.taskPd.dGain[k] = r[k], r[6 + k], r[12 + k], r[18 + k], r[24 + k]


@pytest.mark.parametrize('fp32', [False, True])
def test_task_pd_closed_loop(oracle_mod, fp32):
    import emu_harness as E
    o = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    e = E.EmuSim(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie.cmodel'), fp32=fp32)
    rows = task_rows(np.random.default_rng(2))
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=np.array(PD_PGAIN) * 0.3, dGain=PD_DGAIN)
    fill_task(u, rows)
    e.set_task(rows)
    pd = PD_ROW.copy()
    pd[30:40] *= 0.3
    worst = 0.0
    for k in range(600):
        o.step_pd(u)
        e.step(pd)
        worst = max(worst, np.abs(e.get('qpos')[:35] - o.arr('qpos')).max())
    assert worst < (2e-4 if fp32 else 1e-9), worst
    # the branch really acted: without it the trajectory differs
    o2 = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    u2 = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=np.array(PD_PGAIN) * 0.3, dGain=PD_DGAIN)
    for k in range(600):
        o2.step_pd(u2)
    assert np.abs(o2.arr('qpos') - o.arr('qpos')).max() > 1e-3
    # switching the rows off again restores the motor-only law
    e.set_task(None)
    for k in range(20):
        o.step_pd(u2)
        e.step(pd)
    assert np.abs(e.get('qpos')[:35] - o.arr('qpos')).max() < (2e-4 if fp32 else 1e-9)
