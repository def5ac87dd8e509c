"""GPU: per-environment model constants, the set_const launch and the legacy setter verbs, through the C-ABI, against the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, product
from test_randomise import randomise

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def P():
    return product()


@pytest.mark.parametrize('model', ['cassie', 'cassie_tray_box'])
def test_batch_every_env_its_own_constants(P, oracle_mod, model):
    """4 environments with 4 different sets of constants == 4 oracle sims with the same constants; set_const on a subset"""
    O, n = oracle_mod, 4
    os_ = [O.OracleSim(os.path.join(GOLDEN, model + '.omodel')) for _ in range(n)]
    b = P.CassieBatch(n, modelfile=P.model_path(model), precision=P.FP64)
    nb = os_[0].get_int('nbody')
    rows = {k: [] for k in ('body_mass', 'body_ipos', 'dof_damping', 'geom_friction')}
    base = {k: b.get_model(k) for k in rows}
    for k in rows:                                  # the getters start out with the model file's values
        assert np.abs(base[k][0] - os_[0].model_arr(k)).max() < 1e-15, k
    for i, o in enumerate(os_):
        if i == 0:                                  # env 0 keeps the defaults
            vals = {k: o.model_arr(k).copy() for k in rows}
        else:
            vals = randomise(o, np.random.default_rng(10 + i), free_body=nb - 1 if model == 'cassie_tray_box' else None)
        for k in rows:
            rows[k].append(vals[k])
    for k in rows:
        b.set_model(k, np.array(rows[k]))
        assert np.abs(b.get_model(k) - np.array(rows[k])).max() < 1e-15 or k == 'geom_friction'
    u = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    b.set_pd(P.pd_rows(n, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
    for _ in range(250):
        for o in os_:
            o.step_pd(u)
    b.step(250)
    q = b.qpos()
    for i, o in enumerate(os_):
        assert np.abs(q[i] - o.arr('qpos')).max() < 1e-9, i
    assert np.abs(q[1] - q[2]).max() > 1e-4
    # set_const + state reset for envs 1 and 3 only; 0 and 2 carry on untouched
    mask = np.array([0, 1, 0, 1], dtype=np.uint8)
    before = b.qpos()
    b.set_const(mask, reset_state=True)
    after = b.qpos()
    assert np.array_equal(after[[0, 2]], before[[0, 2]])
    for i in (1, 3):
        os_[i].set_const()
        assert np.abs(after[i] - os_[i].arr('qpos')).max() < 1e-12
    assert np.abs(b.time() - np.array([0.125, 0, 0.125, 0])).max() < 1e-12
    for _ in range(250):
        for o in os_:
            o.step_pd(u)
    b.step(250)
    q = b.qpos()
    for i, o in enumerate(os_):
        assert np.abs(q[i] - o.arr('qpos')).max() < 1e-9, i


def test_rows_off_equals_rows_on_with_defaults(P):
    """the extended kernel instance reading default constant rows follows the plain instance (two separately compiled instances: equal up to
    the compiler's choice of fused multiply-adds, so a tolerance rather than bit equality; the host emulation checks bit equality)"""
    for prec, tol in ((P.FP64, 1e-11), (P.FP32, 1e-5)):
        a, b = P.CassieBatch(3, precision=prec), P.CassieBatch(3, precision=prec)
        rows = P.pd_rows(3, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
        a.set_pd(rows)
        b.set_pd(rows)
        b.set_model('dof_damping', b.get_model('dof_damping'))      # allocates the constant rows, values unchanged
        a.step(300)
        b.step(300)
        assert np.abs(a.qpos() - b.qpos()).max() < tol


def test_free_body_ipos_is_rejected(P):
    b = P.CassieBatch(2, modelfile=P.model_path('cassie_tray_box'), precision=P.FP64)
    ip = b.get_model('body_ipos')
    ip[:, -3:] += 0.01
    with pytest.raises(RuntimeError):
        b.set_model('body_ipos', ip)


def test_legacy_setters(P, oracle_mod):
    """reference verbs on a cassie_sim_t: borrowed model pointers, name-based setters, set_const (src/cassiemujoco.c:1303-1436, 949-977)"""
    O = oracle_mod
    o, c = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel')), P.CassieSim()
    nq, nv, nu, ns, nb, ng = c.params()
    assert (nq, nv, nu, ns, nb, ng) == (35, 32, 10, 29, o.get_int('nbody'), o.get_int('ngeom'))
    assert np.abs(c.get_body_mass() - o.model_arr('body_mass')).max() < 1e-15 and np.abs(c.get_geom_friction().ravel() - o.model_arr('geom_friction')).max() < 1e-15
    m = c.get_body_mass()
    m[1:] *= 1.2
    c.set_body_mass(m)
    o.model_arr('body_mass')[:] = m
    c.set_body_mass(12.5, name='cassie-pelvis')
    o.model_arr('body_mass')[1] = 12.5
    assert c.get_body_mass(name='cassie-pelvis') == 12.5
    d = c.get_dof_damping()
    d[6:] *= 1.5
    c.set_dof_damping(d)
    o.model_arr('dof_damping')[:] = d
    c.set_dof_damping(0.3, name='left-knee')
    o.model_arr('dof_damping')[12] = 0.3
    assert c.get_dof_damping(name='left-knee')[0] == 0.3
    c.set_body_ipos([0.06, 0.0, 0.04], name='cassie-pelvis')
    o.model_arr('body_ipos')[3:6] = [0.06, 0.0, 0.04]
    f = c.get_geom_friction()
    f[:, 0] = 0.6
    c.set_geom_friction(f.ravel())
    o.model_arr('geom_friction')[:] = f.ravel()
    c.set_geom_friction([0.8, 0.005, 0.0001], name='floor')
    o.model_arr('geom_friction')[0:3] = [0.8, 0.005, 0.0001]
    assert np.abs(c.get_geom_friction(name='floor') - [0.8, 0.005, 0.0001]).max() == 0
    c.set_const()
    o.set_const()
    u, pu = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN), P.pd_in_t()
    for side, leg in enumerate((pu.leftLeg, pu.rightLeg)):
        for i in range(5):
            leg.motorPd.pTarget[i] = PD_TARGET[5 * side + i]
            leg.motorPd.pGain[i], leg.motorPd.dGain[i] = PD_PGAIN[i], PD_DGAIN[i]
    for _ in range(400):
        o.step_pd(u)
        c.step_pd(pu)
    assert np.abs(c.qpos() - o.arr('qpos')).max() < 1e-9
    assert abs(c.time() - o.arr('time')[0]) < 1e-12
    # whole-array ipos setter keeps the reference's [i + j] indexing
    ip = np.arange(3 * nb, dtype=float) * 1e-3
    c.set_body_ipos(ip)
    got = c.get_body_ipos().reshape(nb, 3)
    assert np.array_equal(got, np.array([[ip[i + j] for j in range(3)] for i in range(nb)]))
