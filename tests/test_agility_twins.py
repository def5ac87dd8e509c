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
