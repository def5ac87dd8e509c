"""The closed Agility blocks (pd_input_step motor branch, cassie_core_sim_step): twins vs the reference's archive."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REFERENCE, REPO

FUZZ = os.path.join(REPO, 'oracle', '_ref', 'fuzz_agility')
ARCHIVE = os.path.join(REFERENCE, 'src', 'libagilitycassie.a')


def _fuzz_binary():
    if os.path.exists(ARCHIVE):
        subprocess.check_call(['make', '-s', '-C', os.path.join(REPO, 'oracle'), 'fuzz', 'ref'])
    if not os.path.exists(FUZZ):
        pytest.skip('neither the reference archive nor a prebuilt oracle/_ref is available')
    return FUZZ


@pytest.mark.parametrize('mode', [0, 1])
def test_oracle_twins_match_archive_on_fuzz(mode):
    out = subprocess.check_output([_fuzz_binary(), '100000', str(mode)], text=True)
    m = re.search(r'pd twin - archive\| = ([0-9.e+-]+)\s+max\|core twin - archive\| = ([0-9.e+-]+)', out)
    assert m, out
    assert float(m.group(1)) <= 1e-9 and float(m.group(2)) <= 1e-9, out   # torques up to ~1e4 N m with the fuzzed task gains


def test_product_controller_matches_archive_in_closed_loop(oracle_mod):
    """the product's controller stage (step_core.inl, run through the host emulation in fp64) against the oracle linked with the
    REAL archive blocks, over BASELINE config 1 (zero pd_in_t, 1000 ticks: the collapsing robot drives joints into the safety
    layer's soft limits) and the fixed-PD controller"""
    import oracle as O
    _fuzz_binary()
    if not os.path.exists(O.lib_path(ref=True)):
        pytest.skip('oracle/_ref/liboracle_ref.so not available')
    import emu_harness as E
    for u, pd in ((O.make_pd(), np.zeros(50)),
                  (O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN), np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN]))):
        o = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'), ref=True)
        e = E.EmuSim(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie.cmodel'))
        worst, safety = 0.0, 0
        for k in range(1000):
            o.step_pd(u)
            e.step(pd)
            worst = max(worst, np.abs(e.get('qpos')[:35] - o.arr('qpos')).max())
            knee = o.arr('qpos')[14]
            safety += knee < -2.5727
        assert worst < 1e-9, worst
        if not pd.any():
            assert safety > 50     # the safety layer really was exercised


DECODED = ['pelvis.orientation', 'pelvis.rotationalVelocity', 'pelvis.translationalAcceleration', 'leftFoot.position', 'leftFoot.orientation',
           'leftFoot.footRotationalVelocity', 'leftFoot.footTranslationalVelocity', 'rightFoot.position', 'rightFoot.orientation',
           'rightFoot.footRotationalVelocity', 'rightFoot.footTranslationalVelocity', 'motor.position', 'motor.velocity', 'motor.torque',
           'joint.position', 'joint.velocity']


def field(y, path):
    for p in path.split('.'):
        y = getattr(y, p)
    return np.array(y[:])


def test_estimator_twin_matches_archive_in_closed_loop(oracle_mod, pkg):
    """the decoded stateless subset of state_output_step (oracle o_state_output_step) against the REAL estimator, fed the same cassie_out
    along a falling-and-catching trajectory (large joint excursions; an IMU quaternion with w < 0)"""
    import ctypes as C
    import oracle as O
    _fuzz_binary()
    if not os.path.exists(O.lib_path(ref=True)):
        pytest.skip('oracle/_ref/liboracle_ref.so not available')
    L = O.load(ref=True)
    L.o_state_output_step.argtypes = [C.c_void_p, C.c_void_p]
    L.osim_cassie_out.restype = C.c_void_p
    worst = {k: 0.0 for k in DECODED}
    for u, setup in ((O.make_pd(), None), (O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN), 'spin')):
        o = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'), ref=True)
        if setup == 'spin':
            o.arr('qpos')[3:7] = [-0.995, 0.02, -0.03, 0.09]      # same kind of attitude, stored with w < 0: the estimator reports the w >= 0 twin
            o.arr('qvel')[3:6] = [0.5, -0.4, 0.6]
        y, y2, co = pkg.state_out_t(), pkg.state_out_t(), (C.c_char * 1336)()
        saw_neg = False
        for k in range(700):
            o.step_pd(u, y, co)                         # y: real estimator on this tick's cassie_out (copied into co)
            L.o_state_output_step(C.byref(co), C.byref(y2))
            for f in DECODED:
                worst[f] = max(worst[f], np.abs(field(y, f) - field(y2, f)).max())
            saw_neg |= o.arr('sensordata')[16] < 0
        if setup == 'spin':
            assert saw_neg
    for f, w in worst.items():
        assert w < 1e-11, (f, w)


def _probe(O):
    import ctypes as C
    _fuzz_binary()
    if not os.path.exists(O.lib_path(ref=True)):
        pytest.skip('oracle/_ref/liboracle_ref.so not available')
    L = O.load(ref=True)
    L.probe_est.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]

    def est(inp):
        a, out = np.ascontiguousarray(inp, dtype=np.float64), np.zeros(123)
        L.probe_est(a.ctypes.data_as(C.POINTER(C.c_double)), 1, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out
    return est


def test_estimator_twin_on_random_inputs(oracle_mod):
    """o_est_foot against the archive on arbitrary angles / rates (far outside what a trajectory visits: all mat2quat branches)"""
    import ctypes as C
    est, L = _probe(oracle_mod), oracle_mod.load()
    L.o_est_foot.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 6
    rng, worst = np.random.default_rng(11), 0.0
    for it in range(2000):
        x = np.zeros(45)
        x[32] = 1
        x[0:10], x[10:20], x[20:26], x[26:32] = rng.uniform(-3, 3, 10), rng.uniform(-8, 8, 10), rng.uniform(-3, 3, 6), rng.uniform(-8, 8, 6)
        e = est(x)
        for sd in range(2):
            ang = (C.c_double * 7)(*x[5 * sd:5 * sd + 4], x[20 + 3 * sd], x[21 + 3 * sd], x[5 * sd + 4])
            rate = (C.c_double * 7)(*x[10 + 5 * sd:14 + 5 * sd], x[26 + 3 * sd], x[27 + 3 * sd], x[14 + 5 * sd])
            out = [(C.c_double * n)() for n in (3, 4, 3, 3)]
            L.o_est_foot(sd, ang, rate, *out)
            got = np.concatenate([np.array(a[:]) for a in out])
            worst = max(worst, np.abs(got - e[22 + 19 * sd:35 + 19 * sd]).max())
    assert worst < 1e-12, worst


def test_archive_kinematics_pin_the_mjcf_compiler(oracle_mod):
    """an independent check of the model compiler's kinematic tables: the foot point computed by the oracle's mj_kinematics on the compiled
    model/cassie.xml equals the closed estimator's own forward kinematics (which carries Agility's constants) to the MJCF's 5-digit rounding"""
    est = _probe(oracle_mod)
    o = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    names = open(os.path.join(GOLDEN, 'cassie.omodel')).read().split('names_body')[1].split('\n')[0].split()[2:]
    pel, feet = names.index('cassie-pelvis'), (names.index('left-foot'), names.index('right-foot'))
    rng, worst = np.random.default_rng(4), 0.0
    lo, hi = np.array([-0.25, -0.35, -0.8, -2.7, -2.4]), np.array([0.3, 0.35, 1.3, -0.7, -0.6])
    for it in range(50):
        mL, mR, sh, ta = rng.uniform(lo, hi), rng.uniform(lo, hi), rng.uniform(-0.05, 0.05, 2), rng.uniform(0.9, 2.0, 2)
        q = o.arr('qpos')
        q[:] = 0
        q[2], q[3], q[10], q[24] = 1.01, 1, 1, 1
        q[[7, 8, 9, 14, 20]], q[15], q[16] = mL, sh[0], ta[0]
        q[[21, 22, 23, 28, 34]], q[29], q[30] = mR, sh[1], ta[1]
        o.forward()
        xp, xm = o.arr('xpos').reshape(-1, 3), o.arr('xmat').reshape(-1, 3, 3)
        x = np.zeros(45)
        x[32] = 1
        x[0:5], x[5:10], x[20], x[21], x[23], x[24] = mL, mR, sh[0], ta[0], sh[1], ta[1]
        e = est(x)
        for sd, fb in enumerate(feet):
            p = xm[pel].T @ (xp[fb] + xm[fb] @ np.array([0.01762, 0.05219, 0]) - xp[pel])
            worst = max(worst, np.abs(p - e[22 + 19 * sd:25 + 19 * sd]).max())
    assert worst < 5e-7, worst


def test_twins_against_committed_archive_vectors(oracle_mod, pkg):
    """tests/golden/agility_vectors.npz holds outputs of the REAL closed blocks on 300 random inputs (tests/golden/make_golden.py); the oracle's
    twins reproduce them with no archive present -- this is what pins the twins on the GPU box"""
    import ctypes as C
    V = np.load(os.path.join(GOLDEN, 'agility_vectors.npz'))
    L = oracle_mod.load()
    dp = C.POINTER(C.c_double)
    L.o_est_foot.argtypes = [C.c_int] + [dp] * 6
    L.o_pd_input_step.argtypes = [C.c_void_p, C.c_void_p, dp]
    L.o_core_sim_step.argtypes = [dp, C.c_void_p, dp]
    L.o_state_output_step.argtypes = [C.c_void_p, C.c_void_p]
    # build cassie_out_t / pd_in_t images with ctypes mirrors of the bus structs: use the product's mirrors for pd_in_t and a raw
    # buffer + the oracle's own accessor layout for cassie_out_t (offsets from include/cassie_bus.h via a tiny C helper)
    L.osim_fill_cassie_out.argtypes = [C.c_void_p, dp]
    worst = dict(est=0.0, pd=0.0, core=0.0)
    for x, t, u, ch, est, pdq, core in zip(V['cassie_out'], V['task'], V['u'], V['ch8'], V['state_out'], V['pd_torque'], V['core_torque']):
        co = (C.c_char * 1336)()
        xx = np.ascontiguousarray(x)
        L.osim_fill_cassie_out(co, xx.ctypes.data_as(dp))
        y = pkg.state_out_t()
        L.o_state_output_step(co, C.byref(y))
        got = np.concatenate([[0, 0, 0], field(y, 'pelvis.orientation'), field(y, 'pelvis.rotationalVelocity'), [0, 0, 0], field(y, 'pelvis.translationalAcceleration')])
        want = est[:16].copy()
        want[0:3] = 0
        want[10:13] = 0                                  # pelvis position / translational velocity: the filters, checked in tests/test_estimator_filter.py
        worst['est'] = max(worst['est'], np.abs(got - want).max())
        for sd, name in enumerate(('leftFoot', 'rightFoot')):
            g = np.concatenate([field(y, name + '.position'), field(y, name + '.orientation'), field(y, name + '.footRotationalVelocity'), field(y, name + '.footTranslationalVelocity')])
            worst['est'] = max(worst['est'], np.abs(g - est[22 + 19 * sd:35 + 19 * sd]).max())
        pu = pkg.pd_in_t()
        for sd, leg in enumerate((pu.leftLeg, pu.rightLeg)):
            r = t[30 * sd:30 * sd + 30]
            for k in range(6):
                leg.taskPd.torque[k], leg.taskPd.pTarget[k], leg.taskPd.dTarget[k], leg.taskPd.pGain[k], leg.taskPd.dGain[k] = r[k], r[6 + k], r[12 + k], r[18 + k], r[24 + k]
        tq = (C.c_double * 10)()
        L.o_pd_input_step(C.byref(pu), co, tq)
        worst['pd'] = max(worst['pd'], np.abs(np.array(tq[:]) - pdq).max() / max(1.0, np.abs(pdq).max()))
        co2 = (C.c_char * 1336)()
        L.osim_fill_cassie_out(co2, xx.ctypes.data_as(dp))
        L.osim_set_radio8(co2, C.c_double(float(ch)))
        uu, out = np.ascontiguousarray(u), (C.c_double * 10)()
        L.o_core_sim_step(uu.ctypes.data_as(dp), co2, out)
        worst['core'] = max(worst['core'], np.abs(np.array(out[:]) - core).max())
    assert worst['est'] < 1e-11 and worst['pd'] < 1e-12 and worst['core'] < 1e-9, worst


def test_toe_heel_force_twin(oracle_mod, pkg):
    """toeForce / heelForce: the decoded spring-force model against the real estimator.  The archive evaluates it in single precision, so the
    bar is its own rounding noise (a few mN on forces of ~100 N), not 1e-12: |twin - archive| <= 2e-2 N + 2e-4 |F| in closed loop and on the
    committed random vectors whose deflections are physically plausible"""
    import ctypes as C
    import oracle as O
    V = np.load(os.path.join(GOLDEN, 'agility_vectors.npz'))
    L = oracle_mod.load()
    dp = C.POINTER(C.c_double)
    L.o_est_leg_force.argtypes = [C.c_int, dp, dp, dp]
    checked = 0
    for x, est in zip(V['cassie_out'], V['state_out']):
        for sd in range(2):
            m, sh, ta = x[5 * sd:5 * sd + 5], x[20 + 3 * sd], x[21 + 3 * sd]
            if abs(sh) > 0.1 or abs(m[3] + sh + ta - np.deg2rad(13)) > 0.1 or not (-2.5 < m[3] < -0.7) or abs(m[0]) > 0.5 or abs(m[1]) > 0.5 or abs(m[2]) > 1.3:
                continue                                    # outside the leg's working range (the four-bar flips, the x-z solve degenerates)
            ang, q, f = (C.c_double * 7)(m[0], m[1], m[2], m[3], sh, ta, m[4]), (C.c_double * 4)(*x[32:36]), (C.c_double * 3)()
            L.o_est_leg_force(sd, ang, q, f)
            want = est[35 + 19 * sd:38 + 19 * sd]
            assert np.abs(np.array(f[:]) - want).max() <= 2e-2 + 2e-4 * np.abs(want).max(), (sd, f[:], want)
            assert np.array_equal(want, est[38 + 19 * sd:41 + 19 * sd])        # the archive reports the same vector for toe and heel
            checked += 1
    if os.path.exists(O.lib_path(ref=True)):   # closed loop against the live archive
        o = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'), ref=True)
        Lr = O.load(ref=True)
        Lr.o_state_output_step.argtypes = [C.c_void_p, C.c_void_p]
        y, y2, co = pkg.state_out_t(), pkg.state_out_t(), (C.c_char * 1336)()
        u = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
        big = 0.0
        for k in range(800):
            o.step_pd(u, y, co)
            Lr.o_state_output_step(C.byref(co), C.byref(y2))
            for name in ('leftFoot', 'rightFoot'):
                for fld in ('toeForce', 'heelForce'):
                    a, b = field(y, name + '.' + fld), field(y2, name + '.' + fld)
                    assert np.abs(a - b).max() <= 2e-2 + 2e-4 * np.abs(a).max(), (k, name, fld, a, b)
                    big = max(big, np.abs(a).max())
        assert big > 30                                       # standing: the springs carry the weight
        checked += 1
    assert checked > 0
