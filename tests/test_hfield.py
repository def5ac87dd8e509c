"""BASELINE config 4 ingredient: cassie_hfield.xml -- analytic height-field contacts (own definition), product vs oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO, product

OMODEL = os.path.join(GOLDEN, 'cassie_hfield.omodel')
CMODEL = os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie_hfield.cmodel')
PD_ROW = np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN])


def terrain(seed, amp=0.25):
    """test_hfield.c:41-57 style terrain: iid U(0,1) heights (scaled by `amp`), a flat 10x10 patch under the robot"""
    rng = np.random.default_rng(seed)
    d = (rng.uniform(0, 1, (200, 200)) * amp).astype(np.float32)
    d[95:105, 95:105] = 0
    return d


def oracle_with_terrain(O, data):
    o = O.OracleSim(OMODEL)
    C.memmove(o.L.osim_hfield_data(o.h), data.ctypes.data, data.nbytes)
    return o


def test_emulated_kernel_matches_oracle_on_rough_terrain(oracle_mod):
    import emu_harness as E
    data = terrain(7)
    o, e = oracle_with_terrain(oracle_mod, data), E.EmuSim(CMODEL)
    e.set_hfield(data)
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    contacts = 0
    for k in range(1500):
        o.step_pd(u)
        e.step(PD_ROW)
        contacts += o.get_int('ncon')
        if k % 50 == 0 or k == 1499:
            assert np.abs(e.get('qpos')[:35] - o.arr('qpos')).max() < 1e-9, k
            assert int(e.get('counters')[1]) == o.get_int('ncon')
    assert contacts > 1000 and o.arr('qpos')[2] < 0.95        # it really landed on the terrain (surface at z = -0.1 under the feet)


def test_flat_hfield_equals_plane_offset(oracle_mod):
    """an all-zero height field is the plane z = -0.1: same contacts as the plane model lowered by 0.1 m"""
    o = oracle_with_terrain(oracle_mod, np.zeros((200, 200), dtype=np.float32))
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    p = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    p.arr('qpos')[2] += 0.1
    p.forward()
    for _ in range(600):
        o.step_pd(u)
        p.step_pd(u)
    dq = o.arr('qpos') - p.arr('qpos')
    dq[2] += 0.1
    assert np.abs(dq).max() < 1e-6


@pytest.mark.gpu
def test_gpu_hfield_matches_oracle(oracle_mod):
    P = product()
    K, n = 3, 6
    terrains = np.stack([terrain(s) for s in range(K)])
    terrains[1, 95:105, 95:105] = np.random.default_rng(11).uniform(0, 0.1, (10, 10)).astype(np.float32)   # rough (<= 2 cm) under the feet too
    b = P.CassieBatch(n, modelfile=CMODEL, precision=P.FP64)
    b.set_hfield_data(terrains)
    b.set_pd(P.pd_rows(n, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
    oracles = [oracle_with_terrain(oracle_mod, terrains[k]) for k in range(K)]
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    for k in range(1200):
        for o in oracles:
            o.step_pd(u)
        b.step(1)
        if k % 100 == 99:
            q = b.qpos()
            for e in range(n):       # env e stands on terrain e % K
                assert np.abs(q[e] - oracles[e % K].arr('qpos')).max() < 1e-9, (k, e)
    assert np.abs(b.qpos()[0] - b.qpos()[1]).max() > 1e-4     # different terrains, different trajectories
    f = P.CassieBatch(64, modelfile=CMODEL, precision=P.FP32)
    f.set_hfield_data(terrains)
    f.set_pd(P.pd_rows(64, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
    f.step(1200)                      # same horizon as the oracles above
    q = f.qpos()
    assert np.isfinite(q).all() and np.abs(q[0] - oracles[0].arr('qpos')).max() < 5e-3   # fp32 on rough terrain: contact make/break amplifies round-off


@pytest.mark.gpu
def test_gpu_legacy_hfield_verbs():
    P = product()
    sim = P.CassieSim(modelfile=CMODEL)
    L = sim.L
    for n in ('cassie_sim_get_hfield_nrow', 'cassie_sim_get_hfield_ncol', 'cassie_sim_get_nhfielddata'):
        getattr(L, n).argtypes = [C.c_void_p]
    L.cassie_sim_hfielddata.argtypes = [C.c_void_p]
    L.cassie_sim_hfielddata.restype = C.POINTER(C.c_float)
    assert L.cassie_sim_get_hfield_nrow(sim.c) == 200 and L.cassie_sim_get_nhfielddata(sim.c) == 40000
    ptr = L.cassie_sim_hfielddata(sim.c)
    for i in range(40000):
        ptr[i] = 0.5                                   # raise the whole terrain by 0.1 m through the borrowed pointer
    low = P.CassieSim(modelfile=CMODEL)                # same model, terrain left at zero (surface at z = -0.1)
    u = P.pd_in_t()
    for leg, off in ((u.leftLeg, 0), (u.rightLeg, 5)):
        for i in range(5):
            leg.motorPd.pTarget[i] = PD_TARGET[off + i]; leg.motorPd.pGain[i] = PD_PGAIN[i]; leg.motorPd.dGain[i] = PD_DGAIN[i]
    for _ in range(500):
        sim.step_pd(u); low.step_pd(u)
    assert 0.07 < sim.qpos()[2] - low.qpos()[2] < 0.13   # stands 0.1 m higher on the raised surface
