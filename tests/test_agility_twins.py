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
    assert float(m.group(1)) <= 1e-12 and float(m.group(2)) <= 1e-9, out


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
