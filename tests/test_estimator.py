"""The stateless part of the reference's estimator (state_output_step, closed source; decoded in oracle/cassie_oracle.c o_state_output_step
and pinned to the real archive in tests/test_agility_twins.py): the product's observation-stage code against that twin."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO
from test_agility_twins import DECODED, field

PD_ROW = np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN])
OB_EST_ACC, OB_FOOT, OB_EST_QUAT = 56, 60, 86       # devmodel.h


def row_vs_state_out(row, y, tol):
    """observation row of the product against a state_out_t (decoded fields only)"""
    want = np.concatenate([field(y, 'pelvis.translationalAcceleration'), [0]] +
                          [np.concatenate([field(y, s + '.position'), field(y, s + '.orientation'), field(y, s + '.footRotationalVelocity'),
                                           field(y, s + '.footTranslationalVelocity')]) for s in ('leftFoot', 'rightFoot')] + [field(y, 'pelvis.orientation')])
    got = row[OB_EST_ACC:OB_EST_QUAT + 4].copy()
    got[3] = 0
    assert np.abs(got - want).max() < tol, np.abs(got - want).max()
    assert np.abs(row[0:10] - field(y, 'motor.position')).max() < tol and np.abs(row[30:36] - field(y, 'joint.position')).max() < tol


def row_vs_twin(O, row, tol):
    L = O.load()
    L.o_est_foot.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 6
    for sd in range(2):
        mp, mv, jp, jv = row[5 * sd:5 * sd + 5], row[10 + 5 * sd:15 + 5 * sd], row[30 + 3 * sd:33 + 3 * sd], row[36 + 3 * sd:39 + 3 * sd]
        ang = (C.c_double * 7)(mp[0], mp[1], mp[2], mp[3], jp[0], jp[1], mp[4])
        rate = (C.c_double * 7)(mv[0], mv[1], mv[2], mv[3], jv[0], jv[1], mv[4])
        out = [(C.c_double * n)() for n in (3, 4, 3, 3)]
        L.o_est_foot(sd, ang, rate, *out)
        want = np.concatenate([np.array(a[:]) for a in out])
        got = row[OB_FOOT + 13 * sd:OB_FOOT + 13 * sd + 13]
        assert np.abs(got - want).max() < tol * max(1.0, np.abs(want).max()), (sd, np.abs(got - want).max())
    q, w, a = row[42:46], row[46:49], row[49:52]
    R2 = np.array([2 * (q[1] * q[3] - q[0] * q[2]), 2 * (q[2] * q[3] + q[0] * q[1]), 1 - 2 * (q[1] * q[1] + q[2] * q[2])]) / (q @ q)
    r = np.array([0.03155, 0, -0.079996])
    want = a - R2 * 9.806 - np.cross(w, np.cross(w, r))
    eq = row[OB_EST_QUAT:OB_EST_QUAT + 4]
    assert np.abs(row[OB_EST_ACC:OB_EST_ACC + 3] - want).max() < 50 * tol and min(np.abs(eq - q).max(), np.abs(eq + q).max()) < 50 * tol


@pytest.mark.parametrize('fp32', [False, True])
def test_observation_row_matches_estimator_twin(oracle_mod, pkg, fp32):
    import emu_harness as E
    o = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    e = E.EmuSim(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie.cmodel'), fp32=fp32)
    q = o.arr('qpos')
    q[3:7] = [-0.995, 0.02, -0.03, 0.09]        # an attitude stored with w < 0
    o.arr('qvel')[3:6] = [0.5, -0.4, 0.6]
    qe = e.get('qpos')
    qe[:35] = q
    e.set('qpos', qe)
    ve = e.get('qvel')
    ve[3:6] = [0.5, -0.4, 0.6]
    e.set('qvel', ve)
    o.forward()
    e.forward()
    u, y = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN), pkg.state_out_t()
    for k in range(400):
        o.step_pd(u, y)
        e.step(PD_ROW)
        if k % 20 == 19:
            row = e.get('obs')
            if not fp32:
                row_vs_state_out(row, y, 1e-10)
            # self-consistency of the row (also in fp32, where quantised encoders may sit one count away from the fp64 oracle's):
            # the twin applied to the row's OWN measured angles and rates
            row_vs_twin(oracle_mod, row, 2e-5 if fp32 else 1e-12)
    assert y.pelvis.orientation[0] > 0 and o.arr('sensordata')[16] < 0


def test_host_side_leg_force_function(oracle_mod, pkg):
    """the product's host-side toe / heel force (csrc/estimator_host.h, exported as a pure C function: no GPU needed) against the oracle's twin
    (same model, written independently: <= 1e-9 relative) and against the committed outputs of the real archive (its single-precision noise)"""
    L, O = pkg.lib(), oracle_mod.load()
    dp = C.POINTER(C.c_double)
    L.cassie_b200_estimator_leg_force.argtypes = [C.c_int, dp, dp, dp]
    L.cassie_b200_estimator_leg_force.restype = None
    O.o_est_leg_force.argtypes = [C.c_int, dp, dp, dp]
    V = np.load(os.path.join(GOLDEN, 'agility_vectors.npz'))
    n = 0
    for x, est in zip(V['cassie_out'], V['state_out']):
        for sd in range(2):
            m, sh, ta = x[5 * sd:5 * sd + 5], x[20 + 3 * sd], x[21 + 3 * sd]
            if abs(sh) > 0.1 or abs(m[3] + sh + ta - np.deg2rad(13)) > 0.1 or not (-2.5 < m[3] < -0.7) or abs(m[0]) > 0.5 or abs(m[1]) > 0.5 or abs(m[2]) > 1.3:
                continue
            ang, q = (C.c_double * 7)(m[0], m[1], m[2], m[3], sh, ta, m[4]), (C.c_double * 4)(*x[32:36])
            f, g = (C.c_double * 3)(), (C.c_double * 3)()
            L.cassie_b200_estimator_leg_force(sd, ang, q, f)
            O.o_est_leg_force(sd, ang, q, g)
            f, g, want = np.array(f[:]), np.array(g[:]), est[35 + 19 * sd:38 + 19 * sd]
            assert np.abs(f - g).max() <= 1e-9 * max(1.0, np.abs(g).max()), (f, g)
            assert np.abs(f - want).max() <= 2e-2 + 2e-4 * np.abs(want).max(), (f, want)
            n += 1
    assert n > 300
