#!/usr/bin/env python3
"""Regenerates the committed fixtures under tests/golden/ from the reference checkout.
Runs ONLY in the build container (needs /root/reference); the GPU box uses the committed files.

  *.omodel   oracle model tables: oracle/mjcf_compile.py applied to /root/reference/model/<name>.xml
             (derived numeric tables, not a copy of the XML)
  estimator_sequence.npz   400 consecutive 2 kHz calls of the reference's closed estimator (state_output_step) in a closed-loop run from the
             initial state (free fall, touch-down, load transfer): its stateless outputs that feed its filters, and the filters' outputs
  agility_vectors.npz   input/output pairs of the reference's closed Agility blocks (pd_input_step incl. taskPd, cassie_core_sim_step,
             state_output_step from src/libagilitycassie.a via oracle/_ref/liboracle_ref.so; layout: oracle/probe_estimator.c)
"""
import os
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, 'oracle'))
import mjcf_compile as mc  # noqa: E402

REF = os.environ.get('CASSIE_REFERENCE', '/root/reference')
MODELS = ['cassie', 'cassie_hfield', 'cassie_tray_box', 'cassie_no_grav', 'cassie_mass', 'cassie_depth']


def agility_vectors(n=400, seed=123):
    """random cassie_out / pd_in_t inputs -> outputs of the REAL closed blocks (src/libagilitycassie.a through oracle/_ref/liboracle_ref.so)"""
    import ctypes as C
    import subprocess
    import numpy as np
    subprocess.check_call(['make', '-s', '-C', os.path.join(REPO, 'oracle'), 'ref', 'REF=' + REF])
    L = C.CDLL(os.path.join(REPO, 'oracle', '_ref', 'liboracle_ref.so'))
    dp = C.POINTER(C.c_double)
    L.probe_est.argtypes = [dp, C.c_int, dp]
    L.probe_pd.argtypes = [dp, dp, C.c_int, dp]
    L.probe_core.argtypes = [dp, dp, C.c_double, dp]
    rng = np.random.default_rng(seed)
    lo, hi = np.array([-0.3, -0.4, -0.9, -2.9, -2.5] * 2), np.array([0.4, 0.4, 1.4, -0.6, -0.5] * 2)
    X, T, U, CH, EST, PD, CORE = [], [], [], [], [], [], []
    for i in range(n):
        x = np.zeros(45)
        x[0:10] = rng.uniform(lo, hi) if i % 4 else rng.uniform(-3, 3, 10)
        x[10:20] = rng.uniform(-8, 8, 10)
        x[20:26] = [rng.uniform(-0.2, 0.2), rng.uniform(0.8, 2.6), rng.uniform(-2.4, -0.6)] * 2
        if i % 2:   # every other sample inside the leg's working range (small spring deflections): what the toe / heel force model is pinned on
            x[0:10] = rng.uniform(np.array([-0.2, -0.3, -0.6, -2.2, -2.2] * 2), np.array([0.3, 0.3, 1.0, -0.9, -0.8] * 2))
            for sd in range(2):
                x[20 + 3 * sd] = rng.uniform(-0.05, 0.05)
                x[21 + 3 * sd] = np.deg2rad(13) - x[5 * sd + 3] - x[20 + 3 * sd] + rng.uniform(-0.05, 0.05)
        x[26:32] = rng.uniform(-6, 6, 6)
        q = rng.normal(size=4)
        x[32:36] = q / np.linalg.norm(q)
        x[36:39], x[39:42] = rng.uniform(-3, 3, 3), rng.uniform(-12, 12, 3)
        t = np.zeros(60)
        if i % 3:
            t[:] = np.concatenate([np.concatenate([rng.uniform(-30, 30, 6), rng.uniform(-1, 1, 6), rng.uniform(-2, 2, 6), rng.uniform(0, 300, 6), rng.uniform(0, 10, 6)]) for _ in range(2)])
        u, ch = rng.uniform(-250, 250, 10), (1.0 if i % 7 else 0.0)
        e, p, c = np.zeros(123), np.zeros(10), np.zeros(10)
        L.probe_est(x.ctypes.data_as(dp), 1, e.ctypes.data_as(dp))
        L.probe_pd(x.ctypes.data_as(dp), t.ctypes.data_as(dp), 1, p.ctypes.data_as(dp))
        L.probe_core(x.ctypes.data_as(dp), u.ctypes.data_as(dp), ch, c.ctypes.data_as(dp))
        X.append(x); T.append(t); U.append(u); CH.append(ch); EST.append(e[:105]); PD.append(p); CORE.append(c)
    np.savez_compressed(os.path.join(HERE, 'agility_vectors.npz'), cassie_out=np.array(X), task=np.array(T), u=np.array(U), ch8=np.array(CH),
                        state_out=np.array(EST), pd_torque=np.array(PD), core_torque=np.array(CORE))
    print('wrote agility_vectors.npz', n)


def estimator_sequence(T=400):
    """closed loop of the oracle's simulator with the REAL closed blocks (oracle/_ref/liboracle_ref.so): per call the archive's own stateless
    outputs the filters consume (orientation 4, translationalAcceleration 3, foot positions 2x3, toe+heel force 2x3 = 19) and the filter
    outputs (position 3, translationalVelocity 3, externalForce 3, terrain.height = 10)"""
    import ctypes as C
    import importlib
    import numpy as np
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    import oracle as O
    from conftest import PD_DGAIN, PD_PGAIN, PD_TARGET
    pkg = importlib.import_module('cassie-mujoco-sim_b200')
    O.build(ref=True)
    o = O.OracleSim(os.path.join(HERE, 'cassie.omodel'), ref=True)
    u = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    y = pkg.state_out_t()
    X, Y = [], []
    for k in range(T):
        o.step_pd(u, y)
        X.append(np.concatenate([y.pelvis.orientation[:], y.pelvis.translationalAcceleration[:], y.leftFoot.position[:], y.rightFoot.position[:],
                                 np.array(y.leftFoot.toeForce[:]) + np.array(y.leftFoot.heelForce[:]), np.array(y.rightFoot.toeForce[:]) + np.array(y.rightFoot.heelForce[:])]))
        Y.append(np.concatenate([y.pelvis.position[:], y.pelvis.translationalVelocity[:], y.pelvis.externalForce[:], [y.terrain.height]]))
        assert not any(y.pelvis.externalMoment[:]) and not any(y.terrain.slope[:])
    np.savez_compressed(os.path.join(HERE, 'estimator_sequence.npz'), stateless=np.array(X), filtered=np.array(Y))
    print('wrote estimator_sequence.npz', T)


def main():
    agility_vectors()
    estimator_sequence()
    for name in MODELS:
        m = mc.compile_mjcf(os.path.join(REF, 'model', name + '.xml'))
        mc.write_omodel(m, os.path.join(HERE, name + '.omodel'))
        print('wrote', name + '.omodel')


if __name__ == '__main__':
    main()
