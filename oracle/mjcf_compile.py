#!/usr/bin/env python3
"""ORACLE-SIDE MJCF compiler (test infrastructure, NOT product code).

Turns the MJCF subset used by the reference's Cassie models
(/root/reference/model/cassie.xml, cassie_hfield.xml, cassie_tray_box.xml) into the
flat "mjModel-like" constant tables the fp64 C oracle (oracle/cassie_oracle.c) steps on.
It restates what MuJoCo 2.1.0's XML compiler + mj_setConst do for this model family
(the reference obtains these through mj_loadXML, src/cassiemujoco.c:851,997, and
mj_setConst, :952).  MuJoCo itself is an un-vendored binary dependency of the reference
 so this is a restatement of its published behaviour; parity with
MuJoCo is UNPINNED (no MuJoCo in this container).

Deliberately independent of the product's C++ compiler (cassie-mujoco-sim_b200/csrc/mjcf.cpp):
different language, dense numpy linear algebra for the mj_setConst constants (the product
uses the sparse L'DL path).  tests/test_model_compile.py diffs the two.

Usage:  python oracle/mjcf_compile.py /root/reference/model/cassie.xml out.omodel
"""
import sys
import math
import xml.etree.ElementTree as ET
import numpy as np

JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
GEOM_TYPES = {'plane': 0, 'hfield': 1, 'sphere': 2, 'capsule': 3, 'ellipsoid': 4, 'cylinder': 5, 'box': 6, 'mesh': 7}
MINVAL = 1e-15


def fl(s):
    return np.array([float(x) for x in s.split()], dtype=np.float64)


# ---------------------------------------------------------------- quaternion helpers
def quat_mul(a, b):
    return np.array([
        a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3],
        a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2],
        a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1],
        a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0]])


def quat2mat(q):
    w, x, y, z = q
    return np.array([
        [w*w + x*x - y*y - z*z, 2*(x*y - w*z), 2*(x*z + w*y)],
        [2*(x*y + w*z), w*w - x*x + y*y - z*z, 2*(y*z - w*x)],
        [2*(x*z - w*y), 2*(y*z + w*x), w*w - x*x - y*y + z*z]])


def mat2quat(R):
    # numerically robust conversion (largest-diagonal branch)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [0.25*s, (R[2, 1]-R[1, 2])/s, (R[0, 2]-R[2, 0])/s, (R[1, 0]-R[0, 1])/s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1]-R[1, 2])/s, 0.25*s, (R[0, 1]+R[1, 0])/s, (R[0, 2]+R[2, 0])/s]
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2]-R[2, 0])/s, (R[0, 1]+R[1, 0])/s, 0.25*s, (R[1, 2]+R[2, 1])/s]
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0]-R[0, 1])/s, (R[0, 2]+R[2, 0])/s, (R[1, 2]+R[2, 1])/s, 0.25*s]
    q = np.array(q)
    q /= np.linalg.norm(q)
    if q[0] < 0:
        q = -q
    return q


def z2quat(vec):
    """minimal rotation taking +z to vec (MuJoCo 'fromto' convention)."""
    vec = vec / np.linalg.norm(vec)
    axis = np.cross([0, 0, 1.0], vec)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        axis = np.array([1.0, 0, 0])
    else:
        axis = axis / s
    ang = math.atan2(s, vec[2])
    return np.concatenate([[math.cos(ang/2)], axis*math.sin(ang/2)])


def axisangle2quat(axis, ang):
    return np.concatenate([[math.cos(ang/2)], np.asarray(axis)*math.sin(ang/2)])


# ---------------------------------------------------------------- defaults
class Defaults:
    """MJCF default classes: class name -> tag -> attribute dict, parents pre-merged."""

    def __init__(self, root):
        self.cls = {'main': {}}
        top = root.find('default')
        if top is not None:
            self._walk(top, None)

    def _walk(self, node, parent):
        name = node.get('class') or 'main'
        base = {k: dict(v) for k, v in self.cls[parent].items()} if parent else {}
        for child in node:
            if child.tag != 'default':
                base.setdefault(child.tag, {}).update(child.attrib)
        self.cls[name] = base
        for child in node.findall('default'):
            self._walk(child, name)

    def get(self, cls, tag):
        return dict(self.cls[cls or 'main'].get(tag, {}))


def orientation(attrs, degree=True):
    """body/geom/site orientation attributes -> quaternion."""
    if 'quat' in attrs:
        q = fl(attrs['quat'])
        return q / np.linalg.norm(q)
    if 'xyaxes' in attrs:
        v = fl(attrs['xyaxes'])
        x = v[:3] / np.linalg.norm(v[:3])
        y = v[3:] - x * np.dot(x, v[3:])
        y /= np.linalg.norm(y)
        z = np.cross(x, y)
        return mat2quat(np.stack([x, y, z], axis=1))
    if 'zaxis' in attrs:
        return z2quat(fl(attrs['zaxis']))
    if 'euler' in attrs or 'axisangle' in attrs:
        raise NotImplementedError('euler/axisangle not used by the Cassie models')
    return np.array([1.0, 0, 0, 0])


def box_inertia(size, mass):
    sx, sy, sz = size
    return mass / 3.0 * np.array([sy*sy + sz*sz, sx*sx + sz*sz, sx*sx + sy*sy])


def compile_mjcf(path):
    root = ET.parse(path).getroot()
    comp = root.find('compiler')
    degree = (comp.get('angle', 'degree') == 'degree') if comp is not None else True
    ang = math.pi/180 if degree else 1.0
    dfl = Defaults(root)
    opt = root.find('option')
    M = {}
    M['opt_timestep'] = float(opt.get('timestep', 0.002))
    M['opt_gravity'] = fl(opt.get('gravity', '0 0 -9.81'))
    M['opt_iterations'] = int(opt.get('iterations', 100))
    M['opt_tolerance'] = float(opt.get('tolerance', 1e-8))
    M['opt_impratio'] = float(opt.get('impratio', 1))
    M['opt_magnetic'] = fl(opt.get('magnetic', '0 -0.5 0'))
    assert opt.get('solver', 'Newton') == 'PGS' and opt.get('cone', 'pyramidal') == 'pyramidal'

    # hfield assets
    hf = {}
    asset = root.find('asset')
    for h in (asset.findall('hfield') if asset is not None else []):
        hf[h.get('name')] = dict(nrow=int(h.get('nrow')), ncol=int(h.get('ncol')), size=fl(h.get('size')))

    bodies, joints, geoms, sites = [], [], [], []
    bodies.append(dict(name='world', parent=0, pos=np.zeros(3), quat=np.array([1., 0, 0, 0]), ipos=np.zeros(3),
                       iquat=np.array([1., 0, 0, 0]), mass=0.0, inertia=np.zeros(3), joints=[], geoms=[]))

    def add_geom(g, bid, childclass):
        cls = g.get('class', childclass)
        a = dfl.get(cls, 'geom')
        a.update(g.attrib)
        gtype = GEOM_TYPES[a.get('type', 'sphere')]
        if gtype == GEOM_MESH:
            # visual only in every Cassie model (contype = conaffinity = 0); inertia always comes from <inertial>.  Kept in the table
            # (zero size, zero bounding radius) so geom ids and per-geom arrays are numbered as in the reference's model
            assert int(a.get('contype', 1)) == 0 and int(a.get('conaffinity', 1)) == 0
            a = dict(a)
            a.pop('size', None)
            a.pop('fromto', None)
        size = np.zeros(3)
        s = fl(a['size']) if 'size' in a else np.zeros(0)
        pos = fl(a.get('pos', '0 0 0'))
        quat = orientation(a)
        if 'fromto' in a:
            ft = fl(a['fromto'])
            vec = ft[0:3] - ft[3:6]            # MuJoCo: z axis points from 'to' to 'from'
            size[0] = s[0]
            size[1] = np.linalg.norm(vec) / 2
            pos = 0.5 * (ft[0:3] + ft[3:6])
            quat = z2quat(vec)
        else:
            size[:len(s)] = s
        hfid = -1
        if gtype == GEOM_HFIELD:
            hfid = list(hf.keys()).index(a['hfield'])
        if gtype == GEOM_SPHERE:
            rb = size[0]
        elif gtype == GEOM_CAPSULE:
            rb = size[0] + size[1]
        elif gtype == GEOM_BOX:
            rb = float(np.linalg.norm(size))
        elif gtype == GEOM_HFIELD:
            h = list(hf.values())[hfid]['size']
            rb = float(np.linalg.norm([h[0], h[1], max(h[2], h[3])]))
        else:
            rb = 0.0
        geoms.append(dict(name=a.get('name', ''), type=gtype, body=bid, pos=pos, quat=quat, size=size, rbound=rb,
                          contype=int(a.get('contype', 1)), conaffinity=int(a.get('conaffinity', 1)),
                          condim=int(a.get('condim', 3)), priority=int(a.get('priority', 0)),
                          friction=fl(a.get('friction', '1 0.005 0.0001')), solmix=float(a.get('solmix', 1)),
                          solref=fl(a.get('solref', '0.02 1')), solimp=np.concatenate([fl(a.get('solimp', '0.9 0.95 0.001')), [0.5, 2.0]])[:5],
                          margin=float(a.get('margin', 0)), gap=float(a.get('gap', 0)),
                          mass=float(a['mass']) if 'mass' in a else None, hfid=hfid, cls=cls,
                          user=int(float(a.get('user', '0').split()[0])), group=int(a.get('group', 0))))
        bodies[bid]['geoms'].append(len(geoms) - 1)

    def walk(node, parent, childclass):
        for g in node.findall('geom'):
            add_geom(g, parent, childclass)
        for s in node.findall('site'):
            a = dfl.get(s.get('class', childclass), 'site')
            a.update(s.attrib)
            if 'fromto' in a:
                ft = fl(a['fromto'])
                spos, squat = 0.5*(ft[:3]+ft[3:]), z2quat(ft[:3]-ft[3:])
            else:
                spos, squat = fl(a.get('pos', '0 0 0')), orientation(a)
            sites.append(dict(name=a.get('name', ''), body=parent, pos=spos, quat=squat))
        for b in node.findall('body'):
            bid = len(bodies)
            cc = b.get('childclass', childclass)
            bd = dict(name=b.get('name', ''), parent=parent, pos=fl(b.get('pos', '0 0 0')), quat=orientation(b.attrib),
                      joints=[], geoms=[], explicit_inertial=False)
            ine = b.find('inertial')
            if ine is not None:
                bd['explicit_inertial'] = True
                bd['ipos'] = fl(ine.get('pos'))
                bd['mass'] = float(ine.get('mass'))
                if 'fullinertia' in ine.attrib:
                    f = fl(ine.get('fullinertia'))
                    I = np.array([[f[0], f[3], f[4]], [f[3], f[1], f[5]], [f[4], f[5], f[2]]])
                    w, V = np.linalg.eigh(I)
                    order = np.argsort(-w)          # descending, like mju_eig3
                    w, V = w[order], V[:, order]
                    if np.linalg.det(V) < 0:
                        V[:, 2] = -V[:, 2]
                    bd['inertia'] = w
                    bd['iquat'] = mat2quat(V)
                else:
                    bd['inertia'] = fl(ine.get('diaginertia')) if ine.get('diaginertia') else np.zeros(3)   # a point mass (model/cassie_mass.xml:88)
                    bd['iquat'] = orientation(ine.attrib)
            bodies.append(bd)
            for j in list(b.findall('joint')) + list(b.findall('freejoint')):
                a = dfl.get(j.get('class', cc), 'joint')
                a.update(j.attrib)
                if j.tag == 'freejoint':
                    a['type'] = 'free'
                jt = {'free': JNT_FREE, 'ball': JNT_BALL, 'slide': JNT_SLIDE, 'hinge': JNT_HINGE}[a.get('type', 'hinge')]
                limited = a.get('limited', 'false') == 'true' and jt in (JNT_SLIDE, JNT_HINGE, JNT_BALL)
                rng = fl(a.get('range', '0 0'))
                ref = float(a.get('ref', 0))
                sref = float(a.get('springref', 0))
                if jt == JNT_HINGE or jt == JNT_BALL:
                    rng = rng * ang
                if jt == JNT_HINGE:
                    ref *= ang
                    sref *= ang
                if jt == JNT_FREE:
                    limited = False
                axis = fl(a.get('axis', '0 0 1'))
                axis = axis / np.linalg.norm(axis)
                joints.append(dict(name=a.get('name', ''), type=jt, body=bid, pos=fl(a.get('pos', '0 0 0')), axis=axis,
                                   limited=limited, range=rng, ref=ref, springref=sref,
                                   stiffness=float(a.get('stiffness', 0)), damping=float(a.get('damping', 0)),
                                   armature=float(a.get('armature', 0)), margin=float(a.get('margin', 0)),
                                   solref=fl(a.get('solreflimit', '0.02 1')),
                                   solimp=np.concatenate([fl(a.get('solimplimit', '0.9 0.95 0.001')), [0.5, 2.0]])[:5]))
                bd['joints'].append(len(joints) - 1)
            walk(b, bid, cc)

    wb = root.find('worldbody')
    walk(wb, 0, None)

    def ancestors(b):
        out = []
        b = bodies[b]['parent']
        while b > 0:
            out.append(b)
            b = bodies[b]['parent']
        return out

    # bodies without <inertial>: infer from geoms (inertiafromgeom='auto'); only boxes needed (tray, cup_box)
    for bid, bd in enumerate(bodies):
        if bid == 0 or bd.get('explicit_inertial'):
            continue
        if not bd['joints'] and all(not bodies[a]['joints'] for a in ancestors(bid)):
            # static body welded to the world (e.g. the 'floor' body carrying the height field): its inertia never enters the dynamics
            bd['mass'], bd['ipos'], bd['inertia'], bd['iquat'] = 0.0, np.zeros(3), np.zeros(3), np.array([1.0, 0, 0, 0])
            continue
        tot, com, parts = 0.0, np.zeros(3), []
        for gi in bd['geoms']:
            g = geoms[gi]
            assert g['type'] == GEOM_BOX, 'geom-inferred inertia implemented for boxes only'
            vol = 8 * g['size'][0] * g['size'][1] * g['size'][2]
            m = g['mass'] if g['mass'] is not None else 1000.0 * vol
            parts.append((m, g))
            tot += m
            com += m * g['pos']
        com /= tot
        I = np.zeros((3, 3))
        for m, g in parts:
            R = quat2mat(g['quat'])
            Ig = R @ np.diag(box_inertia(g['size'], m)) @ R.T
            d = g['pos'] - com
            I += Ig + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        w, V = np.linalg.eigh(I)
        order = np.argsort(-w)
        w, V = w[order], V[:, order]
        if np.linalg.det(V) < 0:
            V[:, 2] = -V[:, 2]
        bd['mass'], bd['ipos'], bd['inertia'], bd['iquat'] = tot, com, w, mat2quat(V)

    nbody, njnt, ngeom = len(bodies), len(joints), len(geoms)
    # addresses
    qadr, dadr = 0, 0
    for j in joints:
        j['qposadr'], j['dofadr'] = qadr, dadr
        nq_j, nv_j = {JNT_FREE: (7, 6), JNT_BALL: (4, 3), JNT_SLIDE: (1, 1), JNT_HINGE: (1, 1)}[j['type']]
        qadr += nq_j
        dadr += nv_j
    nq, nv = qadr, dadr
    M.update(nq=nq, nv=nv, nbody=nbody, njnt=njnt, ngeom=ngeom, nsite=len(sites))

    body_parent = np.array([b['parent'] for b in bodies], dtype=np.int32)
    body_jntnum = np.array([len(b['joints']) for b in bodies], dtype=np.int32)
    body_jntadr = np.array([b['joints'][0] if b['joints'] else -1 for b in bodies], dtype=np.int32)
    body_dofnum = np.zeros(nbody, dtype=np.int32)
    body_dofadr = -np.ones(nbody, dtype=np.int32)
    dof_body = np.zeros(nv, dtype=np.int32)
    dof_jnt = np.zeros(nv, dtype=np.int32)
    for ji, j in enumerate(joints):
        n = {JNT_FREE: 6, JNT_BALL: 3}.get(j['type'], 1)
        b = j['body']
        if body_dofadr[b] < 0:
            body_dofadr[b] = j['dofadr']
        body_dofnum[b] += n
        dof_body[j['dofadr']:j['dofadr']+n] = b
        dof_jnt[j['dofadr']:j['dofadr']+n] = ji
    # dof parent: previous dof in same body, else last dof of nearest ancestor with dofs
    dof_parent = -np.ones(nv, dtype=np.int32)
    for d in range(nv):
        b = dof_body[d]
        if d > body_dofadr[b]:
            dof_parent[d] = d - 1
        else:
            p = body_parent[b]
            while p > 0 and body_dofnum[p] == 0:
                p = body_parent[p]
            dof_parent[d] = body_dofadr[p] + body_dofnum[p] - 1 if p > 0 else -1
    dof_Madr = np.zeros(nv, dtype=np.int32)
    nM = 0
    for d in range(nv):
        dof_Madr[d] = nM
        k = d
        while k >= 0:
            nM += 1
            k = dof_parent[k]
    body_root = np.zeros(nbody, dtype=np.int32)
    body_weld = np.zeros(nbody, dtype=np.int32)
    for b in range(1, nbody):
        p = body_parent[b]
        body_root[b] = b if p == 0 else body_root[p]
        body_weld[b] = b if body_jntnum[b] > 0 else body_weld[p]

    M.update(nM=nM, body_parentid=body_parent, body_rootid=body_root, body_weldid=body_weld,
             body_jntnum=body_jntnum, body_jntadr=body_jntadr, body_dofnum=body_dofnum, body_dofadr=body_dofadr,
             body_pos=np.array([b['pos'] for b in bodies]), body_quat=np.array([b['quat'] for b in bodies]),
             body_ipos=np.array([b['ipos'] for b in bodies]), body_iquat=np.array([b['iquat'] for b in bodies]),
             body_mass=np.array([b['mass'] for b in bodies]), body_inertia=np.array([b['inertia'] for b in bodies]),
             dof_bodyid=dof_body, dof_jntid=dof_jnt, dof_parentid=dof_parent, dof_Madr=dof_Madr)
    M['jnt_type'] = np.array([j['type'] for j in joints], dtype=np.int32)
    M['jnt_qposadr'] = np.array([j['qposadr'] for j in joints], dtype=np.int32)
    M['jnt_dofadr'] = np.array([j['dofadr'] for j in joints], dtype=np.int32)
    M['jnt_bodyid'] = np.array([j['body'] for j in joints], dtype=np.int32)
    M['jnt_limited'] = np.array([int(j['limited']) for j in joints], dtype=np.int32)
    M['jnt_pos'] = np.array([j['pos'] for j in joints])
    M['jnt_axis'] = np.array([j['axis'] for j in joints])
    M['jnt_stiffness'] = np.array([j['stiffness'] for j in joints])
    M['jnt_range'] = np.array([j['range'] for j in joints])
    M['jnt_margin'] = np.array([j['margin'] for j in joints])
    M['jnt_solref'] = np.array([j['solref'] for j in joints])
    M['jnt_solimp'] = np.array([j['solimp'] for j in joints])
    dof_arm, dof_damp = np.zeros(nv), np.zeros(nv)
    for j in joints:
        n = {JNT_FREE: 6, JNT_BALL: 3}.get(j['type'], 1)
        dof_arm[j['dofadr']:j['dofadr']+n] = j['armature']
        dof_damp[j['dofadr']:j['dofadr']+n] = j['damping']
    M['dof_armature'], M['dof_damping'] = dof_arm, dof_damp

    qpos0, qspring = np.zeros(nq), np.zeros(nq)
    for j in joints:
        a = j['qposadr']
        if j['type'] == JNT_FREE:
            b = bodies[j['body']]
            qpos0[a:a+3] = b['pos']
            qpos0[a+3:a+7] = b['quat']
            qspring[a:a+7] = qpos0[a:a+7]
        elif j['type'] == JNT_BALL:
            qpos0[a] = 1.0
            qspring[a] = 1.0
        else:
            qpos0[a] = j['ref']
            qspring[a] = j['springref']
    M['qpos0'], M['qpos_spring'] = qpos0, qspring

    M['geom_type'] = np.array([g['type'] for g in geoms], dtype=np.int32)
    M['geom_bodyid'] = np.array([g['body'] for g in geoms], dtype=np.int32)
    M['geom_contype'] = np.array([g['contype'] for g in geoms], dtype=np.int32)
    M['geom_conaffinity'] = np.array([g['conaffinity'] for g in geoms], dtype=np.int32)
    M['geom_condim'] = np.array([g['condim'] for g in geoms], dtype=np.int32)
    M['geom_priority'] = np.array([g['priority'] for g in geoms], dtype=np.int32)
    M['geom_hfid'] = np.array([g['hfid'] for g in geoms], dtype=np.int32)
    M['geom_user'] = np.array([g['user'] for g in geoms], dtype=np.int32)     # nuser_geom = 1: 1 obstacle, 2 robot collision geom
    M['geom_group'] = np.array([g['group'] for g in geoms], dtype=np.int32)
    for k in ('pos', 'quat', 'size', 'friction', 'solref', 'solimp'):
        M['geom_' + k] = np.array([g[k] for g in geoms])
    for k in ('rbound', 'solmix', 'margin', 'gap'):
        M['geom_' + k] = np.array([g[k] for g in geoms])
    M['site_bodyid'] = np.array([s['body'] for s in sites], dtype=np.int32)
    M['site_pos'] = np.array([s['pos'] for s in sites]).reshape(-1, 3)
    M['site_quat'] = np.array([s['quat'] for s in sites]).reshape(-1, 4)
    M['names_body'] = [b['name'] for b in bodies]
    M['names_site'] = [s['name'] for s in sites]
    M['names_geom'] = [g['name'] for g in geoms]
    M['names_joint'] = [j['name'] for j in joints]

    nhf = len(hf)
    M['nhfield'] = nhf
    if nhf:
        M['hfield_nrow'] = np.array([h['nrow'] for h in hf.values()], dtype=np.int32)
        M['hfield_ncol'] = np.array([h['ncol'] for h in hf.values()], dtype=np.int32)
        M['hfield_size'] = np.array([h['size'] for h in hf.values()])

    # ---------------- kinematics at qpos0 (needed for connect anchors and mj_setConst)
    def fk(qpos):
        xpos = np.zeros((nbody, 3))
        xquat = np.zeros((nbody, 4))
        xquat[0, 0] = 1
        xanchor = np.zeros((njnt, 3))
        xaxis = np.zeros((njnt, 3))
        for b in range(1, nbody):
            bd = bodies[b]
            if len(bd['joints']) == 1 and joints[bd['joints'][0]]['type'] == JNT_FREE:
                j = joints[bd['joints'][0]]
                a = j['qposadr']
                xpos[b] = qpos[a:a+3]
                q = qpos[a+3:a+7]
                xquat[b] = q / np.linalg.norm(q)
                xanchor[bd['joints'][0]] = xpos[b]
                xaxis[bd['joints'][0]] = j['axis']
                continue
            p = bd['parent']
            Rp = quat2mat(xquat[p])
            pos = xpos[p] + Rp @ bd['pos']
            quat = quat_mul(xquat[p], bd['quat'])
            for ji in bd['joints']:
                j = joints[ji]
                R = quat2mat(quat)
                xaxis[ji] = R @ j['axis']
                xanchor[ji] = pos + R @ j['pos']
                a = j['qposadr']
                if j['type'] == JNT_SLIDE:
                    pos = pos + xaxis[ji] * (qpos[a] - qpos0[a])
                else:
                    if j['type'] == JNT_BALL:
                        ql = qpos[a:a+4] / np.linalg.norm(qpos[a:a+4])
                    else:
                        ql = axisangle2quat(j['axis'], qpos[a] - qpos0[a])
                    quat = quat_mul(quat, ql)
                    pos = xanchor[ji] - quat2mat(quat) @ j['pos']
            xpos[b] = pos
            xquat[b] = quat / np.linalg.norm(quat)
        return xpos, xquat, xanchor, xaxis

    xpos, xquat, xanchor, xaxis = fk(qpos0)
    xmat = np.array([quat2mat(q) for q in xquat])
    xipos = np.array([xpos[b] + xmat[b] @ bodies[b]['ipos'] for b in range(nbody)])

    # ---------------- equality (connect only)
    eqs = []
    eqn = root.find('equality')
    for e in (eqn if eqn is not None else []):
        assert e.tag == 'connect'
        a = dfl.get(e.get('class'), 'equality')
        a.update(e.attrib)
        names = M['names_body']
        b1, b2 = names.index(a['body1']), names.index(a['body2'])
        anchor = fl(a['anchor'])
        gp = xpos[b1] + xmat[b1] @ anchor
        anchor2 = xmat[b2].T @ (gp - xpos[b2])
        eqs.append(dict(obj1=b1, obj2=b2, data=np.concatenate([anchor, anchor2]), solref=fl(a.get('solref', '0.02 1')),
                        solimp=np.concatenate([fl(a.get('solimp', '0.9 0.95 0.001')), [0.5, 2.0]])[:5]))
    M['neq'] = len(eqs)
    M['eq_obj1id'] = np.array([e['obj1'] for e in eqs], dtype=np.int32)
    M['eq_obj2id'] = np.array([e['obj2'] for e in eqs], dtype=np.int32)
    M['eq_data'] = np.array([e['data'] for e in eqs]).reshape(-1, 6)
    M['eq_solref'] = np.array([e['solref'] for e in eqs]).reshape(-1, 2)
    M['eq_solimp'] = np.array([e['solimp'] for e in eqs]).reshape(-1, 5)

    # ---------------- actuators / sensors
    acts = []
    an = root.find('actuator')
    for m_ in (an if an is not None else []):
        a = dfl.get(m_.get('class'), 'motor')
        a.update(m_.attrib)
        ji = M['names_joint'].index(a['joint'])
        acts.append(dict(jnt=ji, gear=float(a.get('gear', '1').split()[0]), ctrlrange=fl(a.get('ctrlrange', '0 0')),
                         ctrllimited=a.get('ctrllimited', 'false') == 'true', user=float(a.get('user', '0').split()[0])))
    M['nu'] = len(acts)
    M['actuator_jntid'] = np.array([a['jnt'] for a in acts], dtype=np.int32)
    M['actuator_gear'] = np.array([a['gear'] for a in acts])
    M['actuator_ctrlrange'] = np.array([a['ctrlrange'] for a in acts]).reshape(-1, 2)
    M['actuator_ctrllimited'] = np.array([int(a['ctrllimited']) for a in acts], dtype=np.int32)
    M['actuator_user'] = np.array([a['user'] for a in acts])
    sens = []
    sn = root.find('sensor')
    actnames = [m_.get('name') for m_ in (an if an is not None else [])]
    for s in (sn if sn is not None else []):
        # type codes private to this table: 0 actuatorpos, 1 jointpos, 2 framequat, 3 gyro, 4 accelerometer, 5 magnetometer
        if s.tag == 'rangefinder':
            continue      # cassie_no_grav.xml: six rangefinders AFTER the 29 numbers the hot path reads (src/cassiemujoco.c:754-773); not modelled
        t = {'actuatorpos': 0, 'jointpos': 1, 'framequat': 2, 'gyro': 3, 'accelerometer': 4, 'magnetometer': 5}[s.tag]
        if t == 0:
            obj = actnames.index(s.get('actuator'))
        elif t == 1:
            obj = M['names_joint'].index(s.get('joint'))
        else:
            obj = M['names_site'].index(s.get('objname') or s.get('site'))
        sens.append(dict(type=t, obj=obj, user=float(s.get('user', 0)), cutoff=float(s.get('cutoff', 0))))
    M['nsensor'] = len(sens)
    M['sensor_type'] = np.array([s['type'] for s in sens], dtype=np.int32)
    M['sensor_objid'] = np.array([s['obj'] for s in sens], dtype=np.int32)
    M['sensor_user'] = np.array([s['user'] for s in sens])
    M['sensor_cutoff'] = np.array([s['cutoff'] for s in sens])

    # ---------------- mj_setConst: dense M at qpos0, invweights, meaninertia, subtree mass
    def jac(point, body):
        jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
        b = body
        while b > 0:
            for ji in reversed(bodies[b]['joints']):
                j = joints[ji]
                d = j['dofadr']
                if j['type'] == JNT_SLIDE:
                    jp[:, d] = xaxis[ji]
                elif j['type'] == JNT_HINGE:
                    jr[:, d] = xaxis[ji]
                    jp[:, d] = np.cross(xaxis[ji], point - xanchor[ji])
                elif j['type'] == JNT_BALL:
                    for k in range(3):
                        ax = xmat[b][:, k]
                        jr[:, d+k] = ax
                        jp[:, d+k] = np.cross(ax, point - xanchor[ji])
                else:
                    for k in range(3):
                        jp[k, d+k] = 1
                        ax = xmat[b][:, k]
                        jr[:, d+3+k] = ax
                        jp[:, d+3+k] = np.cross(ax, point - xpos[b])
            b = body_parent[b]
        return jp, jr

    Mq = np.diag(dof_arm).copy()
    for b in range(1, nbody):
        jp, jr = jac(xipos[b], b)
        Rb = xmat[b] @ quat2mat(bodies[b]['iquat'])
        Iw = Rb @ np.diag(bodies[b]['inertia']) @ Rb.T
        Mq += bodies[b]['mass'] * jp.T @ jp + jr.T @ Iw @ jr
    Minv = np.linalg.inv(Mq)
    binv = np.zeros((nbody, 2))
    for b in range(1, nbody):
        if body_weld[b] == 0:
            continue
        jp, jr = jac(xipos[b], b)
        binv[b, 0] = np.trace(jp @ Minv @ jp.T) / 3
        binv[b, 1] = np.trace(jr @ Minv @ jr.T) / 3
    dinv = np.zeros(nv)
    for j in joints:
        d = j['dofadr']
        if j['type'] in (JNT_SLIDE, JNT_HINGE):
            dinv[d] = Minv[d, d]
        elif j['type'] == JNT_BALL:
            dinv[d:d+3] = np.trace(Minv[d:d+3, d:d+3]) / 3
        else:
            dinv[d:d+3] = np.trace(Minv[d:d+3, d:d+3]) / 3
            dinv[d+3:d+6] = np.trace(Minv[d+3:d+6, d+3:d+6]) / 3
    M['body_invweight0'] = binv
    M['dof_invweight0'] = dinv
    M['stat_meaninertia'] = float(np.trace(Mq) / nv)
    sub = np.array([b['mass'] for b in bodies])
    for b in range(nbody-1, 0, -1):
        sub[body_parent[b]] += sub[b]
    M['body_subtreemass'] = sub
    M['dense_M0'] = Mq       # kept for cross checks (not used by the stepper)
    return M


def write_omodel(M, path):
    with open(path, 'w') as f:
        f.write('# oracle model table generated by oracle/mjcf_compile.py (derived data; see header of that file)\n')
        for k, v in M.items():
            if k == 'dense_M0':
                continue
            if isinstance(v, list):
                f.write('%s S %d %s\n' % (k, len(v), ' '.join(x if x else '-' for x in v)))
            elif isinstance(v, (int, np.integer)):
                f.write('%s I 1 %d\n' % (k, v))
            elif isinstance(v, float):
                f.write('%s F 1 %.17g\n' % (k, v))
            else:
                a = np.asarray(v)
                if a.dtype.kind in 'iu':
                    f.write('%s I %d %s\n' % (k, a.size, ' '.join('%d' % x for x in a.ravel())))
                else:
                    f.write('%s F %d %s\n' % (k, a.size, ' '.join('%.17g' % x for x in a.ravel())))


if __name__ == '__main__':
    m = compile_mjcf(sys.argv[1])
    write_omodel(m, sys.argv[2])
    print('nq %d nv %d nbody %d njnt %d ngeom %d nM %d neq %d nu %d meaninertia %.6g mass %.4f' % (
        m['nq'], m['nv'], m['nbody'], m['njnt'], m['ngeom'], m['nM'], m['neq'], m['nu'], m['stat_meaninertia'],
        m['body_subtreemass'][0]))
