"""Model envelope beyond the three BASELINE models: model/cassie_mass.xml (a 100 kg point mass welded to the pelvis), model/cassie_depth.xml
(15 box obstacles under the floor, rangefinder sites), and the stair boxes of model/cassie.xml:232-246 brought into play with the geom placement verbs
(src/cassiemujoco.c:1466-1541, example/test_terrain.c) -- the product's stepper source in the host emulation against the oracle.  The GPU twins of
these cases are in tests/test_gpu_variants.py."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO

PD_ROW = np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN])
MODELS = os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models')
STAIRS = [('box1', [0.05, 0.135, -0.02], None, [0.25, 0.1, 0.1]),                                   # a 8 cm step under the left foot
          ('box7', [0.0, -0.135, -0.05], [np.cos(0.2), 0, np.sin(0.2), 0], [0.3, 0.1, 0.1])]        # a tilted one under the right foot


def place_stairs(o, e):
    for name, pos, quat, size in STAIRS:
        g = e.set_geom(name, pos=pos, quat=quat, size=size)
        o.model_arr('geom_pos').reshape(-1, 3)[g] = pos
        o.model_arr('geom_size').reshape(-1, 3)[g] = size
        if quat is not None:
            o.model_arr('geom_quat').reshape(-1, 4)[g] = quat
    o.forward()
    e.forward()


@pytest.mark.parametrize('model', ['cassie_mass', 'cassie_depth'])
def test_variant_follows_the_oracle(oracle_mod, model):
    import emu_harness as E
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    for fp32, tol in ((False, 1e-10), (True, 1e-4)):
        o, e = oracle_mod.OracleSim(os.path.join(GOLDEN, model + '.omodel')), E.EmuSim(os.path.join(MODELS, model + '.cmodel'), fp32=fp32)
        e.plain()
        worst = 0.0
        for k in range(500):
            o.step_pd(u)
            e.step(PD_ROW)
            worst = max(worst, np.abs(e.get('qpos')[:35] - o.arr('qpos')).max())
        assert worst < tol, (model, fp32, worst)
    if model == 'cassie_mass':
        o0 = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
        for k in range(500):
            o0.step_pd(u)
        assert o0.arr('qpos')[2] - o.arr('qpos')[2] > 0.05           # 100 kg on the pelvis: the robot squats


def test_stair_boxes_of_cassie_xml(oracle_mod):
    import emu_harness as E
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    for fp32, tol in ((False, 1e-10), (True, 1e-4)):
        o, e = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel')), E.EmuSim(os.path.join(MODELS, 'cassie.cmodel'), fp32=fp32)
        e.plain()
        place_stairs(o, e)
        worst = 0.0
        for k in range(700):
            o.step_pd(u)
            e.step(PD_ROW)
            worst = max(worst, np.abs(e.get('qpos')[:35] - o.arr('qpos')).max())
        assert worst < tol, (fp32, worst)
        pairs = sorted((c['geom1'], c['geom2']) for c in o.contacts())
        assert len(pairs) >= 2 and all(o.model_arr('geom_size').reshape(-1, 3)[g2][0] in (0.25, 0.3) for _, g2 in pairs)   # both feet stand on boxes, not on the floor
        assert int(e.get('counters')[4]) == 0
    # one foot on a box, the other on the floor: floor and box contacts interleave in MuJoCo's body-pair order (left foot's box contact before the
    # right foot's floor contact) although the kernel visits the static-box pairs in a second run -- the Gauss-Seidel sweep is order dependent
    o, e = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel')), E.EmuSim(os.path.join(MODELS, 'cassie.cmodel'))
    name, pos, quat, size = STAIRS[0]
    g = e.set_geom(name, pos=pos, size=size)
    o.model_arr('geom_pos').reshape(-1, 3)[g] = pos
    o.model_arr('geom_size').reshape(-1, 3)[g] = size
    o.forward(); e.forward()
    mixed = False
    for k in range(500):
        o.step_pd(u)
        e.step(PD_ROW)
        gs = [c['geom1'] for c in o.contacts()]
        mixed = mixed or (0 in gs and any(x != 0 for x in gs) and gs != sorted(gs))
        assert int(e.get('counters')[3]) == o.get_int('solver_iter'), k
    assert np.abs(e.get('qpos')[:35] - o.arr('qpos')).max() < 1e-10 and mixed
    # with the boxes parked where the model file puts them nothing changes: the floor carries the robot (all 135 box pairs are candidates, none is near)
    o, e = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel')), E.EmuSim(os.path.join(MODELS, 'cassie.cmodel'))
    for k in range(300):
        o.step_pd(u)
        e.step(PD_ROW)
    assert np.abs(e.get('qpos')[:35] - o.arr('qpos')).max() < 1e-10 and o.get_int('ncon') >= 2 and all(c['geom1'] == 0 for c in o.contacts())   # floor contacts only
