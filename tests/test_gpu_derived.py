"""GPU: derived-quantity rows and the legacy read-only queries (through the C-ABI) against the oracle's restatement of the reference functions."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, product
from test_derived import ANG, CMP, CMV, FF, FPOS, FVEL, HEEL, NCON, OBST, SELF, TOE, check_row

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def P():
    return product()


@pytest.mark.parametrize('model', ['cassie', 'cassie_hfield', 'cassie_tray_box'])
def test_batch_rows_fp64(P, oracle_mod, model):
    O = oracle_mod
    o = O.OracleSim(os.path.join(GOLDEN, model + '.omodel'))
    b = P.CassieBatch(3, modelfile=P.model_path(model), precision=P.FP64)
    with pytest.raises(RuntimeError):
        b.aux()                                   # not enabled yet: loud, not zeros
    b.enable_aux()
    if model == 'cassie_hfield':
        rng = np.random.default_rng(7)
        h = (rng.random((200, 200)) * 0.25).astype(np.float32)
        h[95:105, 95:105] = 0
        np.ctypeslib.as_array(o.L.osim_hfield_data(o.h), shape=(40000,))[:] = h.ravel()
        b.set_hfield_data(h)
    u = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    b.set_pd(P.pd_rows(3, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
    for k in range(600):
        o.step_pd(u)
        b.step(1)
        if k % 40 == 39:
            a = b.aux()
            assert np.array_equal(a[0], a[2])
            check_row(a[1], o)
    assert b.aux()[1][FF + 2] > 50                # standing on something
    # query refreshes only the centre-of-mass slots, for the current state
    before, q0 = b.aux(), b.qpos()
    b.query()
    a = b.aux()
    assert np.abs(a[1][CMP:CMP + 3] - o.cm_position()).max() < 1e-10
    o.forward()
    assert np.abs(a[1][CMV:CMV + 3] - o.cm_velocity()).max() < 1e-9 and np.abs(a[1][ANG:ANG + 3] - o.angular_momentum()).max() < 1e-9
    keep = np.r_[0:CMP, ANG + 3:56]
    assert np.array_equal(a[:, keep], before[:, keep]) and np.array_equal(b.qpos(), q0)


def test_multitick_launch_rows(P, oracle_mod):
    """rows after a 25-tick launch = rows after 25 single ticks"""
    b1, b2 = P.CassieBatch(2, precision=P.FP64), P.CassieBatch(2, precision=P.FP64)
    rows = P.pd_rows(2, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    for b in (b1, b2):
        b.enable_aux()
        b.set_pd(rows)
    for _ in range(12):
        b1.step(25)
        for _ in range(25):
            b2.step(1)
    assert np.array_equal(b1.aux(), b2.aux())


def test_fp32_rows_close(P, oracle_mod):
    O = oracle_mod
    o = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    b = P.CassieBatch(2, precision=P.FP32)
    b.enable_aux()
    u = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    b.set_pd(P.pd_rows(2, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
    for _ in range(400):
        o.step_pd(u)
    b.step(400)
    a, f = b.aux()[0], o.foot_forces()
    assert np.abs(a[FF:FF + 12] - f).max() < 0.02 * np.abs(f).max()
    assert np.abs(a[FPOS:FPOS + 6] - o.foot_positions()).max() < 1e-4
    assert np.abs(a[FVEL:FVEL + 12] - o.foot_velocities()).max() < 2e-2
    assert int(a[NCON]) == o.get_int('ncon')


def test_legacy_queries(P, oracle_mod):
    """the reference's own query verbs on a cassie_sim_t (include/cassiemujoco.h:200-240)"""
    O = oracle_mod
    o = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    c = P.CassieSim()
    u, pu = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN), P.pd_in_t()
    for side, leg in enumerate((pu.leftLeg, pu.rightLeg)):
        for i in range(5):
            leg.motorPd.pTarget[i] = PD_TARGET[5 * side + i]
            leg.motorPd.pGain[i], leg.motorPd.dGain[i] = PD_PGAIN[i], PD_DGAIN[i]
    assert np.abs(np.array(c.foot_pos()) - o.foot_positions()).max() < 1e-12      # right after init (mj_forward)
    for k in range(500):
        o.step_pd(u)
        c.step_pd(pu)
    ff_last = o.foot_forces()
    assert np.abs(c.foot_forces_raw() - ff_last).max() < 1e-8
    t, h = c.get_heeltoe_forces()
    ot, oh = o.heeltoe_forces()
    assert np.abs(t - ot).max() < 1e-8 and np.abs(h - oh).max() < 1e-8
    lf, rf = c.get_foot_forces()
    assert lf > 50 and rf > 50
    assert np.abs(np.array(c.foot_pos()) - o.foot_positions()).max() < 1e-10
    v = np.zeros(12)
    c.foot_vel(v)
    assert np.abs(v - o.foot_velocities()).max() < 1e-8
    assert c.check_self_collision() == o.check_self_collision() and c.check_obstacle_collision() == o.check_obstacle_collision()
    for g in range(3):
        assert c.check_collision(g) == o.geom_collision(g)
    assert np.abs(np.array(c.center_of_mass_position()) - o.cm_position()).max() < 1e-10
    q = c.qpos()
    assert np.abs(q - o.arr('qpos')).max() < 1e-9                                   # the query did not disturb the state
    o.forward()
    assert np.abs(np.array(c.center_of_mass_velocity()) - o.cm_velocity()).max() < 1e-8
    assert np.abs(np.array(c.angular_momentum()) - o.angular_momentum()).max() < 1e-8
    assert np.abs(c.foot_forces_raw() - ff_last).max() < 1e-8                       # the contact group still describes the last step (the reference's
    # cm_* queries re-collide and leave d->contact inconsistent with efc_force; this library keeps the step's values)
    # crossed shins -> self collision after one step from that state
    c2, o2 = P.CassieSim(), O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    q2 = c2.qpos()
    q2[7], q2[21] = -0.3, 0.3
    c2.set_qpos(q2)
    oq = o2.arr('qpos')
    oq[7], oq[21] = -0.3, 0.3
    o2.step_pd(O.make_pd())
    c2.step_pd(P.pd_in_t())
    assert o2.check_self_collision() and c2.check_self_collision()
