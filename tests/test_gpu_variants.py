"""GPU: model envelope beyond the BASELINE models through the C-ABI -- model/cassie_mass.xml, model/cassie_depth.xml, and the stair boxes of
model/cassie.xml placed with the reference's geom verbs (src/cassiemujoco.c:1466-1541) on a cassie_sim_t and with cassie_batch_set_geom_pose on a batch --
against the oracle.  CPU twins (host emulation): tests/test_model_variants.py."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, product
from test_model_variants import STAIRS

pytestmark = pytest.mark.gpu


def _pd(P):
    pu = P.pd_in_t()
    for side, leg in enumerate((pu.leftLeg, pu.rightLeg)):
        for i in range(5):
            leg.motorPd.pTarget[i] = PD_TARGET[5 * side + i]
            leg.motorPd.pGain[i], leg.motorPd.dGain[i] = PD_PGAIN[i], PD_DGAIN[i]
    return pu


@pytest.mark.parametrize('model', ['cassie_mass', 'cassie_depth'])
def test_variant_models(oracle_mod, model):
    P, O = product(), oracle_mod
    u = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    o = O.OracleSim(os.path.join(GOLDEN, model + '.omodel'))
    for _ in range(500):
        o.step_pd(u)
    rows = P.pd_rows(3, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    for prec, tol in ((P.FP64, 1e-9), (P.FP32, 1e-4)):
        b = P.CassieBatch(3, modelfile=P.model_path(model), precision=prec)
        b.set_pd(rows)
        b.step(500)
        assert np.abs(b.qpos()[1] - o.arr('qpos')).max() < tol, (model, prec)
        b.close()


def test_stairs_with_the_legacy_geom_verbs(oracle_mod):
    P, O = product(), oracle_mod
    L = P.lib()
    o, c = O.OracleSim(os.path.join(GOLDEN, 'cassie.omodel')), P.CassieSim()
    dp = __import__('ctypes').POINTER(__import__('ctypes').c_double)
    for fn in ('cassie_sim_geom_name_pos', 'cassie_sim_geom_name_quat', 'cassie_sim_geom_name_size', 'cassie_sim_geom_pos'):
        getattr(L, fn).restype = dp
    for name, pos, quat, size in STAIRS:
        g = None
        for k, nm in enumerate(['box%d' % i for i in range(1, 16)]):
            pass
        arr = (__import__('ctypes').c_double * 3)(*pos)
        L.cassie_sim_set_geom_name_pos(c.c, name.encode(), arr)
        L.cassie_sim_set_geom_name_size(c.c, name.encode(), (__import__('ctypes').c_double * 3)(*size))
        if quat is not None:
            L.cassie_sim_set_geom_name_quat(c.c, name.encode(), (__import__('ctypes').c_double * 4)(*quat))
    # the oracle's model arrays use the same (reference) geom numbering: read the placement back through the borrowed all-geoms pointer
    ng = L.cassie_sim_ngeom(c.c)
    gp = np.array(L.cassie_sim_geom_pos(c.c)[:3 * ng]).reshape(ng, 3)
    L.cassie_sim_geom_quat.restype = dp
    L.cassie_sim_geom_size.restype = dp
    o.model_arr('geom_pos')[:] = gp.ravel()
    o.model_arr('geom_quat')[:] = np.array(L.cassie_sim_geom_quat(c.c)[:4 * ng])
    o.model_arr('geom_size')[:] = np.array(L.cassie_sim_geom_size(c.c)[:3 * ng])
    o.forward()
    u, pu = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN), _pd(P)
    for _ in range(700):
        o.step_pd(u)
        c.step_pd(pu)
    assert np.abs(c.qpos() - o.arr('qpos')).max() < 1e-9
    assert o.arr('qpos')[2] > 0.8 and all(cc['geom1'] != 0 for cc in o.contacts())      # standing on the boxes, the floor plane (geom 0) is untouched
    ff = np.zeros(12)
    L.cassie_sim_foot_forces(c.c, ff.ctypes.data_as(dp))
    assert np.abs(ff - o.foot_forces()).max() < 1e-6 and ff[2] + ff[8] > 150               # the derived-quantity row sees the box contacts: the boxes carry the robot


def test_stairs_in_a_batch():
    P = product()
    n = 64
    a, b = P.CassieBatch(n, precision=P.FP32), P.CassieBatch(n, precision=P.FP32)
    rows = P.pd_rows(n, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    a.set_pd(rows); b.set_pd(rows)
    L = P.lib()
    import ctypes as C
    dp = C.POINTER(C.c_double)
    L.cassie_batch_set_geom_pose.argtypes = [C.c_void_p, C.c_char_p, dp, dp, dp]
    for name, pos, quat, size in STAIRS:
        q = None if quat is None else (C.c_double * 4)(*quat)
        assert L.cassie_batch_set_geom_pose(b.h, name.encode(), (C.c_double * 3)(*pos), q, (C.c_double * 3)(*size)) == 0
    assert L.cassie_batch_set_geom_pose(b.h, b'no-such-geom', None, None, None) == -1
    a.step(600); b.step(600)
    za, zb = a.qpos()[:, 2], b.qpos()[:, 2]
    assert (zb - za > 0.02).all() and np.ptp(zb) < 1e-6                                  # every environment of the batch stands higher, on the boxes
    assert int(b.counters()[:, 4].sum()) == 0
