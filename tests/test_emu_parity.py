"""The product stepper source executed on the host (tests/emu, -DCASSIE_EMU) against the oracle: the no-GPU parity gate."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO

OMODEL = os.path.join(GOLDEN, 'cassie.omodel')
CMODEL = os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie.cmodel')
PD_ROW = np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN])


def test_forward_stages(oracle_mod):
    import emu_harness as E
    D = E.D
    o, e = oracle_mod.OracleSim(OMODEL), E.EmuSim(CMODEL)
    dbg, n = e.get('dbg'), o.get_int('nefc')
    for key, cnt, name in (('XPOS', 78, 'xpos'), ('XQUAT', 104, 'xquat'), ('CDOF', 192, 'cdof'), ('QM', 307, 'qM'), ('QLD', 307, 'qLD'), ('BIAS', 32, 'qfrc_bias'),
                           ('QACCS', 32, 'qacc_smooth'), ('EFC_R', n, 'efc_R'), ('EFC_AREF', n, 'efc_aref'), ('EFC_F', n, 'efc_force'), ('QACC', 32, 'qacc'),
                           ('QFRCC', 32, 'qfrc_constraint'), ('SENS', 29, 'sensordata')):
        want = o.arr(name)[:cnt]
        assert np.abs(dbg[D[key]:D[key] + cnt] - want).max() <= 1e-9 * max(1, np.abs(want).max()), name
    assert np.abs(dbg[D['J']:D['J'] + 32 * n].reshape(n, 32) - o.efc_J()).max() < 1e-13


@pytest.mark.parametrize('cfg', ['zero_pd', 'fixed_pd'])
def test_trajectory_fp64(oracle_mod, cfg):
    import emu_harness as E
    u = oracle_mod.make_pd() if cfg == 'zero_pd' else oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    pd = np.zeros(50) if cfg == 'zero_pd' else PD_ROW
    o, e = oracle_mod.OracleSim(OMODEL), E.EmuSim(CMODEL)
    for k in range(1000):
        o.step_pd(u)
        e.step(pd)
        if k % 20 == 0 or k == 999:
            assert np.abs(e.get('qpos')[:35] - o.arr('qpos')).max() < 1e-10, k
            c = e.get('counters')
            assert int(c[0]) == o.get_int('nefc') and int(c[3]) == o.get_int('solver_iter'), k   # same rows, same PGS iteration count
    assert np.abs(e.get('qvel') - o.arr('qvel')).max() < 1e-8


@pytest.mark.parametrize('cfg', ['zero_pd', 'fixed_pd'])
def test_trajectory_fp32_within_north_star_tolerance(oracle_mod, cfg):
    """fp32 instance of the same source: max|dqpos| <= 1e-4 vs the fp64 oracle over 1000 ticks"""
    import emu_harness as E
    u = oracle_mod.make_pd() if cfg == 'zero_pd' else oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    pd = np.zeros(50) if cfg == 'zero_pd' else PD_ROW
    o, e = oracle_mod.OracleSim(OMODEL), E.EmuSim(CMODEL, fp32=True)
    worst = 0.0
    for k in range(1000):
        o.step_pd(u)
        e.step(pd)
        worst = max(worst, np.abs(e.get('qpos')[:35] - o.arr('qpos')).max())
    assert worst < 1e-4, worst


def test_multitick_and_observation_row(oracle_mod):
    import ctypes as C
    import emu_harness as E
    P = __import__('conftest').product()
    o, a, b = oracle_mod.OracleSim(OMODEL), E.EmuSim(CMODEL), E.EmuSim(CMODEL)
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    y = P.state_out_t()
    for _ in range(30):
        oracle_mod.load().osim_step_pd(o.h, C.byref(u), C.byref(y), None)
        a.step(PD_ROW)
    b.step(PD_ROW, nticks=30)
    assert np.array_equal(a.get('qpos'), b.get('qpos')) and np.array_equal(a.get('obs'), b.get('obs'))
    obs = b.get('obs')
    assert np.abs(obs[0:10] - np.array(y.motor.position)).max() < 1e-12 and np.abs(obs[10:20] - np.array(y.motor.velocity)).max() < 1e-9
    assert np.abs(obs[20:30] - np.array(y.motor.torque)).max() < 1e-9 and np.abs(obs[30:36] - np.array(y.joint.position)).max() < 1e-12
    assert np.abs(obs[42:46] - np.array(y.pelvis.orientation)).max() < 1e-12


def test_torque_delay_and_encoder_quantisation(oracle_mod):
    """a torque step reaches ctrl exactly 6 ticks later (src/cassiemujoco.c:658-661); drive positions sit on the encoder grid"""
    import emu_harness as E
    e = E.EmuSim(CMODEL)
    pd = np.zeros(50)
    pd[1] = 20.0          # left hip-yaw feed-forward torque, output side
    seen = []
    for k in range(9):
        e.step(pd)
        seen.append(e.get('obs')[20 + 1])
    # tick 0 still sees the all-zero cassie_out (calloc, src/cassiemujoco.c:989): knee and foot "positions" of 0 rad violate the safety
    # layer's soft limits by more than 0.15 rad, so every commanded torque is scaled to zero on that tick; the command of tick 1 is the
    # first to pass and reaches the joint 6 ticks later
    assert all(abs(x) < 1e-12 for x in seen[:7]) and abs(seen[7] - 20.0) < 1e-9, seen
    pos = e.get('obs')[0:10]
    bits = [13, 13, 13, 13, 18] * 2
    gear = [25, 25, 16, 16, 50] * 2
    for p, b, g in zip(pos, bits, gear):
        counts = p * g / (2 * np.pi) * (1 << b)
        assert abs(counts - round(counts)) < 1e-6


def test_no_gravity_variant(oracle_mod):
    """model/cassie_no_grav.xml (gravity 0, six rangefinder sensors after the 29 numbers the hot path reads): the robot floats; kernel source
    vs oracle over 500 ticks of PD control"""
    import emu_harness as E
    o = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie_no_grav.omodel'))
    e = E.EmuSim(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie_no_grav.cmodel'))
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    for k in range(500):
        o.step_pd(u)
        e.step(PD_ROW)
    assert np.abs(e.get('qpos')[:35] - o.arr('qpos')).max() < 1e-10
    assert abs(o.arr('qpos')[2] - 1.01) < 0.01 and int(e.get('counters')[1]) == 0      # still floating, no contacts


def test_in_kernel_estimator_matches_oracle(oracle_mod, pkg):
    """the estimator stage of the extended instance (leg forces + filters, every tick of a launch) against the oracle's restated estimator:
    fp64 to rounding, also when several ticks run per launch"""
    import os
    import numpy as np
    from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET
    from emu_harness import EmuSim
    O = oracle_mod

    def rows(y):
        return np.concatenate([y.pelvis.position[:], y.pelvis.translationalVelocity[:], y.pelvis.externalForce[:], [y.terrain.height], y.leftFoot.toeForce[:], y.rightFoot.toeForce[:]])
    u = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    row = np.zeros(50)
    row[10:20], row[30:40], row[40:50] = PD_TARGET, PD_PGAIN, PD_DGAIN
    for nt in (1, 4):
        o, e, y = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel')), EmuSim(pkg.model_path('cassie')), pkg.state_out_t()
        e.enable_est()
        worst = 0.0
        for k in range(0, 600, nt):
            for _ in range(nt):
                o.step_pd(u, y)
            e.step(row, nt)
            a = rows(y)
            worst = max(worst, (np.abs(a - e.get('est_out')) / (1 + np.abs(a))).max())
        assert worst < 1e-10 and abs(a[8]) > 1 and a[12] < -50, (nt, worst)
        e.close()


def test_open_loop_gait_follows_per_tick_targets(oracle_mod):
    """cassie_batch_set_pd_gait (BASELINE config 5's "random PD gaits"): the kernel moves pTarget itself every control tick of a multi-tick
    launch; the oracle is handed the same targets from outside, one pd_in_t per tick, as a caller of the reference would"""
    import emu_harness as E
    amp, ph, f = np.array([0.05, 0.05, 0.3, 0.4, 0.3] * 2), np.array([0.3] * 5 + [0.3 + np.pi] * 5), 1.2
    o, e = oracle_mod.OracleSim(OMODEL), E.EmuSim(CMODEL)
    e.plain()
    e.set_gait(amp, ph, f)
    for k in range(450):
        o.step_pd(oracle_mod.make_pd(pTarget=np.array(PD_TARGET) + amp * np.sin(2 * np.pi * f * k * 0.0005 + ph), pGain=PD_PGAIN, dGain=PD_DGAIN))
        if k % 3 == 0:
            e.step(PD_ROW, 3)
        if k % 3 == 2:
            assert np.abs(e.get('qpos')[:35] - o.arr('qpos')).max() < 1e-10, k
    assert abs(o.arr('qpos')[14] - (-1.1997)) > 0.05      # the knees really moved
    e.set_gait(None, None, None)
    q = e.get('qpos').copy()
    e.step(PD_ROW, 1)
    assert np.abs(e.get('qpos') - q).max() > 0


def test_table_free_shortcuts_change_nothing_beyond_rounding(pkg):
    """The stepper takes model-specific shortcuts that must not change the mathematics: the pelvis' subtree sums composed from its children's sums
    (CASSIE_B200_NOKIDS switches them off) and the code generated for the Cassie dof tree (CASSIE_B200_NOSPEC).  The environment variables are read
    when a model is built, so the same emulation library runs all three variants; 300 ticks of the standing controller, fp64."""
    import subprocess, sys
    code = (
        "import sys, importlib, numpy as np; sys.path.insert(0, %r); import emu_harness as E; from conftest import PD_TARGET, PD_PGAIN, PD_DGAIN\n"
        "P = importlib.import_module('cassie-mujoco-sim_b200')\n"
        "pd = np.zeros(50); pd[10:20] = PD_TARGET; pd[30:40] = (list(PD_PGAIN) * 2)[:10]; pd[40:50] = (list(PD_DGAIN) * 2)[:10]\n"
        "s = E.EmuSim(P.model_path('cassie')); s.step(pd, 300); print(' '.join(repr(float(x)) for x in s.get('qpos', 35)))\n" % os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for extra in ({}, {'CASSIE_B200_NOKIDS': '1'}, {'CASSIE_B200_NOSPEC': '1'}):
        env = dict(os.environ); env.update(extra)
        r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.array([float(x) for x in r.stdout.strip().split()]))
    assert outs[0].size == 35 and np.isfinite(outs[0]).all() and outs[0][2] > 0.8
    for o in outs[1:]:
        d = np.abs(o - outs[0]).max()
        assert d < 1e-11, d
