"""Product MJCF compiler (C++, csrc/mjcf.cpp) vs the oracle-side compiler (Python, oracle/mjcf_compile.py) vs committed tables."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE, REPO, have_reference

RENAME = {'opt_timestep': 'timestep', 'opt_gravity': 'gravity', 'opt_magnetic': 'magnetic', 'opt_tolerance': 'tolerance', 'opt_impratio': 'impratio',
          'opt_iterations': 'iterations', 'stat_meaninertia': 'meaninertia'}


def load_table(path):
    d = {}
    for line in open(path):
        if line[0] == '#':
            continue
        t = line.split()
        k, ty, n = RENAME.get(t[0], t[0]), t[1], int(t[2])
        d[k] = t[3:3 + n] if ty == 'S' else np.array([float(x) for x in t[3:3 + n]])
    return d


def q2m(q):
    w, x, y, z = q
    return np.array([[w*w+x*x-y*y-z*z, 2*(x*y-w*z), 2*(x*z+w*y)], [2*(x*y+w*z), w*w-x*x+y*y-z*z, 2*(y*z-w*x)], [2*(x*z-w*y), 2*(y*z+w*x), w*w-x*x-y*y+z*z]])


def compare(o, c):
    skip = {'body_iquat', 'hfield_nrow', 'hfield_ncol', 'hfield_size'}
    for k in sorted(set(o) & set(c)):
        if k in skip:
            continue
        if isinstance(o[k], list):
            assert o[k] == c[k], k
            continue
        assert o[k].shape == c[k].shape, k
        if o[k].size:
            scale = max(1.0, np.abs(o[k]).max())
            assert np.abs(o[k] - c[k]).max() <= 1e-9 * scale, (k, np.abs(o[k] - c[k]).max())
    for must in ('body_pos', 'body_inertia', 'dof_invweight0', 'body_invweight0', 'eq_data', 'geom_size', 'qpos0', 'meaninertia', 'dof_Madr'):
        assert must in o and must in c, must
    nb = len(o['body_mass'])
    for b in range(nb):   # principal-axis sign/order conventions differ; the world inertia tensor must not
        Io = q2m(o['body_iquat'][4*b:4*b+4]) @ np.diag(o['body_inertia'][3*b:3*b+3]) @ q2m(o['body_iquat'][4*b:4*b+4]).T
        Ic = q2m(c['body_iquat'][4*b:4*b+4]) @ np.diag(c['body_inertia'][3*b:3*b+3]) @ q2m(c['body_iquat'][4*b:4*b+4]).T
        assert np.abs(Io - Ic).max() < 1e-12


def test_committed_tables_agree():
    """tests/golden/cassie.omodel (oracle compiler) vs cassie-mujoco-sim_b200/models/cassie.cmodel (product compiler)"""
    compare(load_table(os.path.join(GOLDEN, 'cassie.omodel')), load_table(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie.cmodel')))


@pytest.mark.skipif(not have_reference(), reason='reference checkout not present')
def test_fresh_compiles_from_reference_xml_agree(tmp_path):
    xml = os.path.join(REFERENCE, 'model', 'cassie.xml')
    sys.path.insert(0, os.path.join(REPO, 'oracle'))
    import mjcf_compile as mc
    m = mc.compile_mjcf(xml)
    mc.write_omodel(m, str(tmp_path / 'a.omodel'))
    subprocess.check_call([sys.executable, os.path.join(REPO, 'tools', 'compile_model.py'), xml, str(tmp_path / 'a.cmodel')])
    o, c = load_table(str(tmp_path / 'a.omodel')), load_table(str(tmp_path / 'a.cmodel'))
    compare(o, c)
    # and the committed fixtures are what the current compilers produce
    compare(o, load_table(os.path.join(GOLDEN, 'cassie.omodel')))
    compare(c, load_table(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie.cmodel')))
    # model facts from SURVEY.md section 8 / Appendix A
    assert (int(o['nq'][0]), int(o['nv'][0]), int(o['nbody'][0]), int(o['njnt'][0]), int(o['nM'][0]), int(o['neq'][0])) == (35, 32, 26, 26, 307, 4)
    assert [int(x) for x in o['dof_parentid']] == [-1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 8, 12, 13, 14, 14, 16, 14, 5, 19, 20, 21, 22, 23, 21, 25, 26, 27, 27, 29, 27]
    assert abs(o['body_subtreemass'][0] - 33.312) < 2e-3


def test_cmodel_roundtrip(tmp_path):
    """load_cmodel(save_cmodel(m)) is the identity on the numbers"""
    src = os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie.cmodel')
    subprocess.check_call([sys.executable, os.path.join(REPO, 'tools', 'compile_model.py'), src, str(tmp_path / 'b.cmodel')])
    a, b = load_table(src), load_table(str(tmp_path / 'b.cmodel'))
    for k in a:
        if isinstance(a[k], list):
            assert a[k] == b[k]
        else:
            assert np.array_equal(a[k], b[k]), k
