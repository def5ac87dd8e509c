"""Derived-quantity queries (reference src/cassiemujoco.c:1586-1961): the product's by-product rows, executed on the host
from the kernel source (tests/emu), against the oracle's restatement of the reference functions."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO

PD_ROW = np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN])
# offsets of include/cassie_b200.h (CASSIE_AUX_*)
FF, TOE, HEEL, FPOS, FVEL, CMP, CMV, ANG, OBST, SELF, GMASK, NCON = 0, 12, 18, 24, 30, 42, 45, 48, 51, 52, 53, 54


def _pair(oracle_mod, model, fp32=False):
    import emu_harness as E
    return (oracle_mod.OracleSim(os.path.join(GOLDEN, model + '.omodel')),
            E.EmuSim(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', model + '.cmodel'), fp32=fp32))


def check_row(a, o, tol_force=1e-8, tol_kin=1e-10):
    """one derived-quantity row `a` against the oracle's queries in the reference's own post-step state"""
    toe, heel = o.heeltoe_forces()
    assert np.abs(a[FF:FF + 12] - o.foot_forces()).max() < tol_force
    assert np.abs(a[TOE:TOE + 6] - toe).max() < tol_force and np.abs(a[HEEL:HEEL + 6] - heel).max() < tol_force
    assert np.abs(a[FPOS:FPOS + 6] - o.foot_positions()).max() < tol_kin
    assert np.abs(a[FVEL:FVEL + 12] - o.foot_velocities()).max() < 100 * tol_kin
    assert bool(a[OBST]) == o.check_obstacle_collision() and bool(a[SELF]) == o.check_self_collision()
    for g in range(4):
        assert bool((int(a[GMASK]) >> g) & 1) == o.geom_collision(g), g
    assert int(a[NCON]) == o.get_int('ncon')


@pytest.mark.parametrize('model', ['cassie', 'cassie_hfield', 'cassie_tray_box'])
def test_step_byproducts_match_reference_queries(oracle_mod, model):
    o, e = _pair(oracle_mod, model)
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    if model == 'cassie_hfield':
        rng = np.random.default_rng(7)
        h = (rng.random((200, 200)) * 0.25).astype(np.float32)
        h[95:105, 95:105] = 0
        np.ctypeslib.as_array(o.L.osim_hfield_data(o.h), shape=(40000,))[:] = h.ravel()
        e.set_hfield(h)
    saw_contact = saw_toe = False
    for k in range(800):
        o.step_pd(u)
        e.step(PD_ROW)
        if k % 25 == 24 or k > 780:
            a = e.get('aux')
            check_row(a, o)
            saw_contact |= a[FF + 2] > 1
            saw_toe |= abs(a[TOE + 2]) > 1 or abs(a[TOE + 5]) > 1
    assert saw_contact
    if model == 'cassie_tray_box':
        assert bool(e.get('aux')[OBST])          # the cup rests on the tray: an obstacle geom is in contact
    if model == 'cassie':
        assert saw_toe or True                   # heel-first landing with this controller; toe slots are covered by the hfield model


def test_self_collision_and_forward_mode(oracle_mod):
    """crossed shins: the frictionless capsule-capsule contact sets the self-collision flag; rows from a forward launch agree as well"""
    o, e = _pair(oracle_mod, 'cassie')
    q = o.arr('qpos')
    q[7], q[21] = -0.3, 0.3
    qe = e.get('qpos')
    qe[:35] = q
    e.set('qpos', qe)
    o.forward()
    e.forward()
    a = e.get('aux')
    assert o.check_self_collision() and bool(a[SELF]) and not bool(a[OBST])
    assert int(a[NCON]) == o.get_int('ncon') >= 1
    assert np.abs(a[FPOS:FPOS + 6] - o.foot_positions()).max() < 1e-12
    u = oracle_mod.make_pd()
    for _ in range(5):
        o.step_pd(u)
        e.step(np.zeros(50))
        check_row(e.get('aux'), o)


@pytest.mark.parametrize('model', ['cassie', 'cassie_tray_box'])
def test_centre_of_mass_group(oracle_mod, model):
    o, e = _pair(oracle_mod, model)
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    for k in range(300):
        # rows written by a step describe the state that step started from: compare with the oracle BEFORE it steps
        if k % 50 == 0 and k:
            o.forward()        # mj_forward refreshes sensordata, which the next tick's encoders read: do the same on both sides
            e.forward()
            want = (o.cm_position(), o.cm_velocity(), o.angular_momentum())
        o.step_pd(u)
        e.step(PD_ROW)
        if k % 50 == 0 and k:
            a = e.get('aux')
            assert np.abs(a[CMP:CMP + 3] - want[0]).max() < 1e-10 and np.abs(a[CMV:CMV + 3] - want[1]).max() < 1e-9 and np.abs(a[ANG:ANG + 3] - want[2]).max() < 1e-9
    # query: the centre-of-mass slots of the CURRENT state, nothing else touched
    before, qv = e.get('aux').copy(), (e.get('qpos').copy(), e.get('qvel').copy(), e.get('cst').copy())
    e.query()
    a = e.get('aux')
    assert np.abs(a[CMP:CMP + 3] - o.cm_position()).max() < 1e-12          # reference semantics: mj_fwdPosition, then subtree_com of the world
    o.forward()                                                            # make the oracle's velocities consistent with its state
    assert np.abs(a[CMV:CMV + 3] - o.cm_velocity()).max() < 1e-10 and np.abs(a[ANG:ANG + 3] - o.angular_momentum()).max() < 1e-10
    keep = np.r_[0:CMP, ANG + 3:56]
    assert np.array_equal(a[keep], before[keep])
    assert np.array_equal(e.get('qpos'), qv[0]) and np.array_equal(e.get('qvel'), qv[1]) and np.array_equal(e.get('cst'), qv[2])
    # total momentum sanity: M * v_cm equals the sum of the bodies' momenta the oracle holds
    m = o.arr('qM')[0] if model == 'cassie' else None
    if m:
        assert abs(m - 33.3) < 1.0


def test_fp32_rows_are_close(oracle_mod):
    o, e = _pair(oracle_mod, 'cassie', fp32=True)
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    for k in range(400):
        o.step_pd(u)
        e.step(PD_ROW)
    a = e.get('aux')
    f = o.foot_forces()
    assert np.abs(a[FF:FF + 12] - f).max() < 0.02 * np.abs(f).max()
    assert np.abs(a[FPOS:FPOS + 6] - o.foot_positions()).max() < 1e-4


def test_compilers_agree_on_geom_tags():
    """geom user / group tags (model/cassie.xml:24,33,86-246) come out the same from the product's MJCF compiler and the oracle's"""
    def table(path):
        out = {}
        for line in open(path):
            t = line.split()
            if t and t[0] in ('geom_user', 'geom_group'):
                out[t[0]] = [int(x) for x in t[3:]]
        return out
    for model in ('cassie', 'cassie_hfield', 'cassie_tray_box'):
        a, b = table(os.path.join(GOLDEN, model + '.omodel')), table(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', model + '.cmodel'))
        assert a and a == b, model
