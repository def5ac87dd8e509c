"""Per-environment model constants + mj_setConst (reference setters src/cassiemujoco.c:1303-1436, set_const :949-977):
the kernel source on the host (tests/emu) against the oracle with the same constants written into its private model."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO

PD_ROW = np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN])
CE_BINVW, CE_DINVW, CE_ROOT, CE_TOT, CE_PGS = 192, 224, 256, 257, 258     # devmodel.h CE_*


def randomise(o, rng, free_body=None):
    """writes randomised constants into the oracle's model arrays and returns them (reference numbering)"""
    nb, ng = o.get_int('nbody'), o.get_int('ngeom')
    mass, ipos, damp, fr = o.model_arr('body_mass'), o.model_arr('body_ipos'), o.model_arr('dof_damping'), o.model_arr('geom_friction')
    mass[1:] *= rng.uniform(0.7, 1.3, nb - 1)
    for b in range(1, nb):
        if b != free_body:
            ipos[3 * b:3 * b + 3] += rng.uniform(-0.01, 0.01, 3)
    damp[:32] *= rng.uniform(0.5, 2.0, 32)
    fr[0::3] *= rng.uniform(0.5, 1.2, ng)
    return dict(body_mass=mass.copy(), body_ipos=ipos.copy(), dof_damping=damp.copy(), geom_friction=fr.copy())


@pytest.mark.parametrize('model', ['cassie', 'cassie_hfield', 'cassie_tray_box'])
def test_randomised_constants_and_set_const(oracle_mod, model):
    import emu_harness as E
    o = oracle_mod.OracleSim(os.path.join(GOLDEN, model + '.omodel'))
    e = E.EmuSim(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', model + '.cmodel'))
    nb, nv = o.get_int('nbody'), o.get_int('nv')
    vals = randomise(o, np.random.default_rng(3), free_body=nb - 1 if model == 'cassie_tray_box' else None)
    for k, v in vals.items():
        e.model_set(k, v)
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    # (1) no set_const yet: the new masses / damping / friction act at once, the solver's reference weights are stale on both sides
    for k in range(300):
        o.step_pd(u)
        e.step(PD_ROW)
    assert np.abs(e.get('qpos')[:o.nq] - o.arr('qpos')).max() < 1e-10
    # the constants really changed the motion
    o0 = oracle_mod.OracleSim(os.path.join(GOLDEN, model + '.omodel'))
    for k in range(300):
        o0.step_pd(u)
    assert np.abs(o0.arr('qpos') - o.arr('qpos')).max() > 1e-4
    # (2) mj_setConst on both sides
    o.set_const()
    e.set_const()
    c = e.get('cenv')
    assert np.abs(c[CE_BINVW:CE_BINVW + nb] - o.model_arr('body_invweight0')[0::2]).max() < 1e-10
    assert np.abs(c[CE_DINVW:CE_DINVW + 32] / o.model_arr('dof_invweight0')[:32] - 1).max() < 1e-10
    assert abs(c[CE_PGS] * o.model_arr('meaninertia')[0] * nv - 1) < 1e-12
    assert abs(1 / c[CE_TOT] - o.model_arr('body_subtreemass')[0]) < 1e-10
    # cassie_sim_set_const also resets qpos / qvel / time and forwards (:955-971); the emulation exposes only the mode-3 launch
    q = e.get('qpos')
    q[:o.nq] = o.arr('qpos')
    e.set('qpos', q)
    e.set('qvel', np.zeros(32))
    if model == 'cassie_tray_box':
        e.set('xqvel', np.zeros(6))
    cst = e.get('cst')
    cst[186] = 0
    e.set('cst', cst)
    e.forward()
    for k in range(300):
        o.step_pd(u)
        e.step(PD_ROW)
    assert np.abs(e.get('qpos')[:o.nq] - o.arr('qpos')).max() < 1e-10


def test_default_rows_reproduce_the_shared_model(oracle_mod):
    """switching the constant row on without changing anything must not change a single bit"""
    import emu_harness as E
    path = os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie.cmodel')
    a, b = E.EmuSim(path), E.EmuSim(path)
    a.plain()          # a: the plain kernel instance; b: the extended instance reading its (default) constant row
    b.enable_cenv()
    for k in range(200):
        a.step(PD_ROW)
        b.step(PD_ROW)
    assert np.array_equal(a.get('qpos'), b.get('qpos')) and np.array_equal(a.get('qvel'), b.get('qvel'))
    b.set_const()      # recomputed weights differ from the compiler's dense ones only by rounding
    c = b.get('cenv')
    o = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    assert np.abs(c[CE_DINVW:CE_DINVW + 32] / o.model_arr('dof_invweight0') - 1).max() < 1e-9


def test_oracle_set_const_matches_independent_dense_compile(oracle_mod):
    """the oracle's C mj_setConst (sparse solves) against the Python compiler's dense inverse that produced the golden tables"""
    o = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie_tray_box.omodel'))
    b0, d0, m0 = o.model_arr('body_invweight0').copy(), o.model_arr('dof_invweight0').copy(), o.model_arr('meaninertia').copy()
    o.just_set_const()
    assert np.abs(o.model_arr('body_invweight0') / np.where(b0 == 0, 1, b0) - np.where(b0 == 0, 0, 1)).max() < 1e-9
    assert np.abs(o.model_arr('dof_invweight0') / d0 - 1).max() < 1e-9 and abs(o.model_arr('meaninertia')[0] / m0[0] - 1) < 1e-12
