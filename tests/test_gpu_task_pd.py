"""GPU: taskPd through the C-ABI (legacy verb, batched AoS, compact rows) against the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, product
from test_task_pd import fill_task, task_rows

pytestmark = pytest.mark.gpu


def test_task_pd_all_entry_points(oracle_mod):
    P, O = product(), oracle_mod
    rows = task_rows(np.random.default_rng(2))
    o = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    u = O.make_pd(pTarget=PD_TARGET, pGain=np.array(PD_PGAIN) * 0.3, dGain=PD_DGAIN)
    fill_task(u, rows)
    pu = P.pd_in_t()
    for side, leg in enumerate((pu.leftLeg, pu.rightLeg)):
        for i in range(5):
            leg.motorPd.pTarget[i], leg.motorPd.pGain[i], leg.motorPd.dGain[i] = PD_TARGET[5 * side + i], PD_PGAIN[i] * 0.3, PD_DGAIN[i]
    fill_task(pu, rows)
    c = P.CassieSim()
    b = P.CassieBatch(3, precision=P.FP64)
    b.set_pd(P.pd_rows(3, pTarget=PD_TARGET, pGain=np.array(PD_PGAIN) * 0.3, dGain=PD_DGAIN))
    b.set_task_pd(np.tile(rows, (3, 1)))
    a = P.CassieBatch(2, precision=P.FP64)
    pin = (P.pd_in_t * 2)(pu, P.pd_in_t())              # env 1 of the AoS batch runs with an all-zero pd_in_t
    o0 = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    for k in range(400):
        o.step_pd(u)
        o0.step_pd(O.make_pd())
        c.step_pd(pu)
        a.step_pd(pin, want_state=False)
    b.step(400)
    assert np.abs(c.qpos() - o.arr('qpos')).max() < 1e-9
    assert np.abs(b.qpos() - o.arr('qpos')).max() < 1e-9
    qa = a.qpos()
    assert np.abs(qa[0] - o.arr('qpos')).max() < 1e-9 and np.abs(qa[1] - o0.arr('qpos')).max() < 1e-9
    b.set_task_pd(None)                                 # off again: same as a batch that never had it
    d = P.CassieBatch(3, precision=P.FP64)
    d.set_qpos(b.qpos())
    d.set_qvel(b.qvel())


def test_task_entry_found_in_a_later_chunk():
    """cassie_sim_step_pd_batch scans for taskPd entries while it packs each chunk of a large batch: an entry that first shows up in the second chunk
    installs the task rows for that chunk and the later ones (the first chunk has already been launched without them).  Against the compact path."""
    P = product()
    n, who = 2085, 1800                                   # two chunks (1056 + 1029); only environment 1800 uses the task branch
    rows = task_rows(np.random.default_rng(5))
    pd = (P.pd_in_t * n)()
    for e in range(n):
        for side, leg in enumerate((pd[e].leftLeg, pd[e].rightLeg)):
            for i in range(5):
                leg.motorPd.pTarget[i], leg.motorPd.pGain[i], leg.motorPd.dGain[i] = PD_TARGET[5 * side + i], PD_PGAIN[i] * 0.3, PD_DGAIN[i]
    fill_task(pd[who], rows)
    a, b = P.CassieBatch(n, precision=P.FP64), P.CassieBatch(n, precision=P.FP64)
    b.set_pd(P.pd_rows(n, pTarget=PD_TARGET, pGain=np.array(PD_PGAIN) * 0.3, dGain=PD_DGAIN))
    tr = np.zeros((n, 60)); tr[who] = rows
    b.set_task_pd(tr)
    for _ in range(40):
        a.step_pd(pd)
        b.step(1)
    qa, qb = a.qpos(), b.qpos()
    assert np.abs(qa - qb).max() < 1e-9
    assert np.abs(qa[who] - qa[who - 1]).max() > 1e-4     # the task controller did act on that environment only
    assert np.abs(qa[0] - qa[who - 1]).max() < 1e-12
