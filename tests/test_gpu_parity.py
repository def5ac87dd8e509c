"""GPU parity: the CUDA stepper (through the C-ABI) against the fp64 oracle on the same inputs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, product

pytestmark = pytest.mark.gpu
OMODEL = os.path.join(GOLDEN, 'cassie.omodel')


@pytest.fixture(scope='module')
def P():
    os.environ['CASSIE_B200_DEBUG'] = '1'
    return product()


def _oracle_traj(O, u, nt, setup=None):
    o = O.OracleSim(OMODEL)
    if setup:
        setup(o)
    qs, vs = [], []
    for _ in range(nt):
        o.step_pd(u)
        qs.append(o.arr('qpos').copy())
        vs.append(o.arr('qvel').copy())
    return o, np.array(qs), np.array(vs)


def test_stage_dump_fp64_matches_oracle(P, oracle_mod):
    """every pipeline stage of mj_forward at the initial state, fp64 kernel vs oracle"""
    from emu_harness import D
    O = oracle_mod
    o = O.OracleSim(OMODEL)
    b = P.CassieBatch(2, precision=P.FP64)
    dbg = b.debug_dump(1)
    assert dbg is not None
    n = o.get_int('nefc')
    pairs = [('XPOS', 78, 'xpos'), ('XQUAT', 104, 'xquat'), ('CDOF', 192, 'cdof'), ('QM', 307, 'qM'), ('QLD', 307, 'qLD'), ('BIAS', 32, 'qfrc_bias'),
             ('SMOOTH', 32, 'qfrc_smooth'), ('QACCS', 32, 'qacc_smooth'), ('EFC_R', n, 'efc_R'), ('EFC_AREF', n, 'efc_aref'), ('EFC_B', n, 'efc_b'),
             ('EFC_F', n, 'efc_force'), ('QACC', 32, 'qacc'), ('QFRCC', 32, 'qfrc_constraint'), ('SENS', 29, 'sensordata')]
    assert int(dbg[D['COUNTS']]) == n
    for key, cnt, name in pairs:
        got, want = dbg[D[key]:D[key] + cnt], o.arr(name)[:cnt]
        tol = 1e-9 * max(1.0, np.abs(want).max())
        assert np.abs(got - want).max() <= tol, (name, np.abs(got - want).max())
    J = dbg[D['J']:D['J'] + 32 * n].reshape(n, 32)
    assert np.abs(J - o.efc_J()).max() < 1e-12


@pytest.mark.parametrize('cfg', ['zero_pd', 'fixed_pd'])
def test_trajectory_fp64(P, oracle_mod, cfg):
    """BASELINE config 1 (zero pd_in_t) and config 2 controller, 1000 ticks, fp64 kernel vs oracle: <= 1e-9 on qpos"""
    O = oracle_mod
    if cfg == 'zero_pd':
        u, rows = O.make_pd(), P.pd_rows(3)
    else:
        u, rows = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN), P.pd_rows(3, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    _, qs, vs = _oracle_traj(O, u, 1000)
    b = P.CassieBatch(3, precision=P.FP64)
    b.set_pd(rows)
    worst = 0.0
    for k in range(1000):
        b.step(1)
        if k % 50 == 49 or k < 3:
            q = b.qpos()
            assert np.abs(q[0] - q[2]).max() == 0.0          # identical envs stay bit-identical
            worst = max(worst, np.abs(q[1] - qs[k]).max())
    assert worst < 1e-9, worst
    assert np.abs(b.qvel()[1] - vs[-1]).max() < 1e-7
    assert abs(b.time()[0] - 0.5) < 1e-12


@pytest.mark.parametrize('cfg', ['zero_pd', 'fixed_pd'])
def test_trajectory_fp32_tolerance(P, oracle_mod, cfg):
    """fp32 throughput build vs fp64 oracle: max|dqpos| <= 1e-4 over 1000 ticks (north_star tolerance)"""
    O = oracle_mod
    if cfg == 'zero_pd':
        u, rows = O.make_pd(), P.pd_rows(2)
    else:
        u, rows = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN), P.pd_rows(2, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    _, qs, _ = _oracle_traj(O, u, 1000)
    b = P.CassieBatch(2, precision=P.FP32)
    b.set_pd(rows)
    worst = 0.0
    for k in range(0, 1000, 25):
        b.step(25)                                         # multi-tick launches must equal single-tick ones
        worst = max(worst, np.abs(b.qpos()[0] - qs[k + 24]).max())
    assert worst < 1e-4, worst


def test_multitick_equals_singletick(P):
    rows = P.pd_rows(4, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    a, b = P.CassieBatch(4, precision=P.FP64), P.CassieBatch(4, precision=P.FP64)
    a.set_pd(rows); b.set_pd(rows)
    for _ in range(40):
        a.step(1)
    b.step(40)
    assert np.array_equal(a.qpos(), b.qpos()) and np.array_equal(a.qvel(), b.qvel())


def test_pelvis_push_matches_oracle(P, oracle_mod):
    """BASELINE config 3 ingredient: xfrc_applied on cassie-pelvis (cassie_sim_apply_force)"""
    O = oracle_mod
    u = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    o = O.OracleSim(OMODEL)
    push = np.array([60.0, -35.0, 0, 0, 0, 0])
    b = P.CassieBatch(2, precision=P.FP64)
    b.set_pd(P.pd_rows(2, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
    for k in range(300):
        if k == 100:
            o.arr('xfrc_applied').reshape(-1, 6)[1] = push
            assert b.apply_force(np.stack([push, np.zeros(6)]), 'cassie-pelvis') == 0
        if k == 200:
            o.arr('xfrc_applied')[:] = 0
            b.clear_forces()
        o.step_pd(u); b.step(1)
    q = b.qpos()
    assert np.abs(q[0] - o.arr('qpos')).max() < 1e-9
    assert np.abs(q[0] - q[1]).max() > 1e-3      # the un-pushed env went elsewhere
    assert b.apply_force(push, 'no-such-body') == -1


def test_aos_entry_point_and_legacy_sim(P, oracle_mod):
    """cassie_sim_step_pd_batch (AoS pd_in_t[] / state_out_t[]) and the legacy cassie_sim_step_pd on a batch of one"""
    O = oracle_mod
    n = 3
    pd = (P.pd_in_t * n)()
    for e in range(n):
        for i in range(5):
            for leg, off in ((pd[e].leftLeg, 0), (pd[e].rightLeg, 5)):
                leg.motorPd.pTarget[i] = PD_TARGET[off + i]; leg.motorPd.pGain[i] = PD_PGAIN[i]; leg.motorPd.dGain[i] = PD_DGAIN[i]
    u = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    o = O.OracleSim(OMODEL)
    b = P.CassieBatch(n, precision=P.FP64)
    sim = P.CassieSim()
    yo = None
    for k in range(60):
        y = b.step_pd(pd)
        ys = sim.step_pd(pd[0])
        import ctypes as C
        yo = P.state_out_t()
        o.L.osim_step_pd(o.h, C.byref(u), C.byref(yo), None)
    for fld in ('position', 'velocity', 'torque'):
        got, leg, want = np.array(getattr(y[1].motor, fld)), np.array(getattr(ys.motor, fld)), np.array(getattr(yo.motor, fld))
        assert np.abs(got - want).max() < 1e-9 and np.abs(leg - want).max() < 1e-9, fld
    assert np.abs(np.array(y[2].joint.position) - np.array(yo.joint.position)).max() < 1e-9
    assert np.abs(np.array(y[0].pelvis.orientation) - np.array(yo.pelvis.orientation)).max() < 1e-9
    assert np.abs(sim.qpos() - o.arr('qpos')).max() < 1e-9 and abs(sim.time() - 0.03) < 1e-12
    # borrowed-pointer write-through: move the pelvis up, step, and see it fall from there
    q = sim.qpos(); q[2] += 0.5; sim.set_qpos(q)
    sim.step_pd(pd[0])
    assert abs(sim.qpos()[2] - q[2]) < 1e-2


def test_batch_diversity_and_reset(P):
    """different PD rows per env give different trajectories; masked reset restores exactly the init state"""
    n = 64
    rng = np.random.default_rng(0)
    rows = P.pd_rows(n, pTarget=np.array(PD_TARGET) + rng.uniform(-0.05, 0.05, (n, 10)), pGain=PD_PGAIN, dGain=PD_DGAIN)
    b = P.CassieBatch(n, precision=P.FP32)
    q0 = b.qpos()
    b.set_pd(rows); b.step(200)
    q = b.qpos()
    assert np.isfinite(q).all() and np.unique(np.round(q[:, 9], 5)).size > n // 2
    quat = q[:, 3:7]
    assert np.abs(np.linalg.norm(quat, axis=1) - 1).max() < 1e-5
    mask = np.zeros(n, dtype=np.uint8); mask[::2] = 1
    b.reset(mask)
    q2 = b.qpos()
    assert np.array_equal(q2[::2], q0[::2]) and np.array_equal(q2[1::2], q[1::2])
    c = b.counters()
    assert (c[:, 0] >= 12).all() and (c[:, 4] == 0).all()


@pytest.mark.parametrize('prec', ['fp64', 'fp32'])
def test_integrate_pos_kernel(P, prec):
    """cassie_batch_integrate_pos == mj_integratePos on random state (quaternion joints included); both instances of the TMA-pipelined kernel -- the fp32
    one is the instance bench.py times against the HBM roofline (tolerance: fp32 rounding of O(1) numbers, 2e-6), over more than one tile per CTA"""
    n = 257 if prec == 'fp64' else 40001
    rng = np.random.default_rng(1)
    b = P.CassieBatch(n, precision=P.FP64 if prec == 'fp64' else P.FP32)
    q = b.qpos(); v = rng.normal(size=(n, 32))
    b.set_qvel(v); b.integrate_pos(); b.sync()
    q1 = b.qpos()
    h = 5e-4
    want = q.copy()
    qa_ball = {3: 3, 10: 9, 24: 22}      # qpos adr -> dof adr of the three ball joints; everything else: dof index = qpos index - (number of preceding ball joints)
    qi, di = 0, 0
    while qi < 35:
        if qi in qa_ball:
            w = v[:, di:di + 3]; nrm = np.linalg.norm(w, axis=1, keepdims=True); ang = h * nrm; ax = w / nrm
            c = np.concatenate([np.cos(ang / 2), ax * np.sin(ang / 2)], axis=1); a = q[:, qi:qi + 4] / np.linalg.norm(q[:, qi:qi + 4], axis=1, keepdims=True)
            want[:, qi:qi + 4] = np.stack([a[:, 0]*c[:, 0]-a[:, 1]*c[:, 1]-a[:, 2]*c[:, 2]-a[:, 3]*c[:, 3], a[:, 0]*c[:, 1]+a[:, 1]*c[:, 0]+a[:, 2]*c[:, 3]-a[:, 3]*c[:, 2],
                                           a[:, 0]*c[:, 2]-a[:, 1]*c[:, 3]+a[:, 2]*c[:, 0]+a[:, 3]*c[:, 1], a[:, 0]*c[:, 3]+a[:, 1]*c[:, 2]-a[:, 2]*c[:, 1]+a[:, 3]*c[:, 0]], axis=1)
            qi += 4; di += 3
        else:
            want[:, qi] = q[:, qi] + h * v[:, di]; qi += 1; di += 1
    assert np.abs(q1 - want).max() < (1e-12 if prec == 'fp64' else 2e-6)


@pytest.mark.parametrize('n', [2085, 5000])
def test_aos_chunked_launches_equal_one_launch(P, n):
    """cassie_sim_step_pd_batch steps large batches as several launches on their own streams (host pack / unpack overlapped with the kernels); the result
    must be, bit for bit, what one launch over the compact rows gives.  2085: two ragged halves; 5000: more than two rounds, chunks of whole rounds."""
    rng = np.random.default_rng(n)
    tgt = np.array(PD_TARGET) + rng.uniform(-0.05, 0.05, (n, 10))
    pd = (P.pd_in_t * n)()
    for e in range(n):
        for i in range(5):
            for leg, off in ((pd[e].leftLeg, 0), (pd[e].rightLeg, 5)):
                leg.motorPd.pTarget[i] = tgt[e, off + i]; leg.motorPd.pGain[i] = PD_PGAIN[i]; leg.motorPd.dGain[i] = PD_DGAIN[i]
    a, b = P.CassieBatch(n, precision=P.FP32), P.CassieBatch(n, precision=P.FP32)
    b.enable_estimator_device(True)      # the AoS entry point switches the in-kernel estimator on by itself: same kernel instance on both sides
    b.set_pd(P.pd_rows(n, pTarget=tgt, pGain=PD_PGAIN, dGain=PD_DGAIN))
    for _ in range(25):
        y = a.step_pd(pd)
        b.step(1)
    assert np.array_equal(a.qpos(), b.qpos()) and np.array_equal(a.qvel(), b.qvel())
    ob = b.obs()
    for e in (0, n // 2 - 1, n // 2, n // 2 + 1, n - 1, 2071, 2072):
        assert np.array_equal(np.array(y[e].motor.position, dtype=np.float32), ob[e, P.OBS['motor_pos']].astype(np.float32)), e
        assert np.array_equal(np.array(y[e].pelvis.orientation, dtype=np.float32), ob[e, P.OBS['est_quat']].astype(np.float32)), e
    assert np.unique(np.round(a.qpos()[:, 9], 6)).size > n // 4
