"""GPU: the legacy verbs RL wrappers call around cassie_sim_step_pd -- state snapshots (cassie_get_state / cassie_set_state, cassie_sim_copy /
duplicate, src/cassiemujoco.c:1072-1093, 3380-3452), cassie_sim_step (:1137-1145), cassie_sim_step_pd_no2khz (:1159-1181), the timestep
accessors (:1196-1204), hold / release (:1974-2000) -- against the oracle's restatement of the same functions, and the batched snapshot verbs."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, product

pytestmark = pytest.mark.gpu


def _pd(P, scale=1.0):
    pu = P.pd_in_t()
    for side, leg in enumerate((pu.leftLeg, pu.rightLeg)):
        for i in range(5):
            leg.motorPd.pTarget[i] = PD_TARGET[5 * side + i] * scale
            leg.motorPd.pGain[i], leg.motorPd.dGain[i] = PD_PGAIN[i], PD_DGAIN[i]
    return pu


def _y(y):
    return np.concatenate([y.pelvis.position[:], y.pelvis.translationalVelocity[:], y.pelvis.externalForce[:], [y.terrain.height], y.motor.position[:], y.motor.velocity[:],
                           y.motor.torque[:], y.joint.velocity[:], y.leftFoot.toeForce[:], y.rightFoot.position[:]])


def test_state_snapshot_restores_everything_bitwise():
    P = product()
    c, pu = P.CassieSim(), _pd(P)
    for _ in range(120):
        c.step_pd(pu)
    s = c.get_state()
    assert np.array_equal(s.qpos(), c.qpos()) and s.time() == pytest.approx(0.06)
    a = [_y(c.step_pd(pu)) for _ in range(60)]
    qa = c.qpos()
    c.set_state(s)
    assert c.time() == pytest.approx(0.06)
    b = [_y(c.step_pd(pu)) for _ in range(60)]
    # encoder filters, torque delay line, warm start, estimator filters: any row left out of the snapshot would show up here
    assert np.array_equal(np.array(a), np.array(b)) and np.array_equal(qa, c.qpos())
    # writes through the state's borrowed pointers are honoured (reference: cassie_state_qpos returns mjData's own array)
    q = s.qpos(); q[2] += 0.05
    s.set_qpos(q); s.set_time(1.5)
    c.set_state(s)
    assert c.qpos()[2] == pytest.approx(q[2]) and c.time() == 1.5
    # a duplicate carries model and state; both continue identically
    c.set_state(c.get_state(s))
    d = c.duplicate()
    ya, yb = [_y(c.step_pd(pu)) for _ in range(40)], [_y(d.step_pd(pu)) for _ in range(40)]
    assert np.array_equal(np.array(ya), np.array(yb)) and np.array_equal(c.qpos(), d.qpos())
    s2 = P.CassieState()
    P.lib().cassie_state_copy(s2.s, s.s)
    assert np.array_equal(s2.qpos(), s.qpos())


def test_torque_level_step_and_cassie_out(oracle_mod):
    """cassie_sim_step(user torques) = pd_input_step reduced to the identity; cassie_out_t fields against the oracle's bus"""
    P, O = product(), oracle_mod
    o, c = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel')), P.CassieSim()
    tq = [3.0, -2.0, 10.0, 30.0, -4.0, -3.0, 2.0, 10.0, 30.0, -4.0]
    u, ui = O.make_pd(torque=tq), P.cassie_user_in_t()
    for i in range(10):
        ui.torque[i] = tq[i]
    co = P.cassie_out_t()
    for k in range(200):
        o.step_pd(u, None, co)
        yc = c.step(ui)
    assert np.abs(c.qpos() - o.arr('qpos')).max() < 1e-9
    for leg in ('leftLeg', 'rightLeg'):
        for dr in ('hipRollDrive', 'hipYawDrive', 'hipPitchDrive', 'kneeDrive', 'footDrive'):
            a, b = getattr(getattr(co, leg), dr), getattr(getattr(yc, leg), dr)
            assert abs(a.position - b.position) < 1e-9 and abs(a.velocity - b.velocity) < 1e-7 and abs(a.torque - b.torque) < 1e-7 and a.torqueLimit == b.torqueLimit and a.gearRatio == b.gearRatio
        for jn in ('shinJoint', 'tarsusJoint', 'footJoint'):
            a, b = getattr(getattr(co, leg), jn), getattr(getattr(yc, leg), jn)
            assert abs(a.position - b.position) < 1e-9 and abs(a.velocity - b.velocity) < 1e-6
    assert np.abs(np.array(co.pelvis.vectorNav.orientation[:]) - np.array(yc.pelvis.vectorNav.orientation[:])).max() < 1e-9
    assert np.abs(np.array(co.pelvis.vectorNav.linearAcceleration[:]) - np.array(yc.pelvis.vectorNav.linearAcceleration[:])).max() < 1e-6
    assert yc.isCalibrated and yc.pelvis.radio.channel[8] == 1 and yc.pelvis.battery.voltage[3] == 4.2 and yc.leftLeg.kneeDrive.statusWord == 0x0637
    g = c.get_cassie_out()
    assert g.leftLeg.kneeDrive.position == yc.leftLeg.kneeDrive.position


def test_timestep_and_no2khz(oracle_mod):
    P, O = product(), oracle_mod
    o, c = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel')), P.CassieSim()
    u, pu = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN), _pd(P)
    assert c.timestep() == 0.0005
    c.set_timestep(0.00025)                       # two physics sub-steps per 2 kHz control tick from now on
    o.model_arr('timestep')[0] = 0.00025
    for k in range(150):
        o.step_pd(u)
        c.step_pd(pu)
    assert c.time() == pytest.approx(150 * 0.0005) and np.abs(c.qpos() - o.arr('qpos')).max() < 1e-9
    for k in range(100):                          # one physics step per call, controller every call
        o.step_pd_no2khz(u)
        c.step_pd_no2khz(pu)
    assert c.time() == pytest.approx(150 * 0.0005 + 100 * 0.00025) and np.abs(c.qpos() - o.arr('qpos')).max() < 1e-9
    P.lib().cassie_sim_timestep(c.c)[0] = 0.0005   # through the borrowed pointer
    o.model_arr('timestep')[0] = 0.0005
    for k in range(50):
        o.step_pd(u)
        c.step_pd(pu)
    assert np.abs(c.qpos() - o.arr('qpos')).max() < 1e-9


def test_hold_and_release(oracle_mod):
    P, O = product(), oracle_mod
    o, c = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel')), P.CassieSim()
    u, pu = O.make_pd(), P.pd_in_t()
    q0 = c.qpos()
    c.hold()
    o.model_arr('jnt_stiffness')[:3] = 1e5; o.model_arr('dof_damping')[:6] = 1e4; o.model_arr('qpos_spring')[:3] = o.arr('qpos')[:3]
    for _ in range(300):
        o.step_pd(u); c.step_pd(pu)
    assert np.abs(c.qpos()[:3] - q0[:3]).max() < 5e-3 and np.abs(c.qpos() - o.arr('qpos')).max() < 1e-8    # hangs in the air, legs sag
    c.release()
    o.model_arr('jnt_stiffness')[:3] = 0; o.model_arr('dof_damping')[:6] = 0
    for _ in range(300):
        o.step_pd(u); c.step_pd(pu)
    assert c.qpos()[2] < q0[2] - 0.02 and np.abs(c.qpos() - o.arr('qpos')).max() < 1e-7


def test_batched_snapshot_masked_restore():
    P = product()
    n = 6
    b = P.CassieBatch(n, precision=P.FP32)
    rows = P.pd_rows(n, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    b.set_pd(rows)
    b.enable_estimator_device()
    b.step(300)
    snap = b.get_state()
    q0, o0 = b.qpos(), b.obs()
    b.apply_force(np.tile([300.0, 0, 0, 0, 0, 0], (n, 1)))      # everybody is pushed over ...
    b.step(400)
    assert (np.abs(b.qpos() - q0).max(axis=1) > 0.05).all()
    fallen = b.qpos()
    m = np.array([1, 0, 1, 0, 0, 1], dtype=np.uint8)
    b.set_state(snap, m)                                         # ... and three of them go back to the stored standing state
    q1 = b.qpos()
    assert np.array_equal(q1[m == 1], q0[m == 1]) and np.array_equal(q1[m == 0], fallen[m == 0])
    assert np.array_equal(b.obs()[m == 1], o0[m == 1])
    b.clear_forces()
    b.set_state(snap)
    b.step(50)
    r = P.CassieBatch(n, precision=P.FP32)
    r.set_pd(rows); r.enable_estimator_device(); r.step(350)
    assert np.array_equal(b.qpos(), r.qpos()) and np.array_equal(b.obs(), r.obs())     # restored + 50 ticks == never disturbed
    b.free_state(snap)
