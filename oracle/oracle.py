"""ctypes harness around the CPU oracle (oracle/cassie_oracle.c).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


class PdMotorIn(C.Structure):
    _fields_ = [(n, C.c_double * 5) for n in ('torque', 'pTarget', 'dTarget', 'pGain', 'dGain')]


class PdTaskIn(C.Structure):
    _fields_ = [(n, C.c_double * 6) for n in ('torque', 'pTarget', 'dTarget', 'pGain', 'dGain')]


class PdLegIn(C.Structure):
    _fields_ = [('taskPd', PdTaskIn), ('motorPd', PdMotorIn)]


class PdIn(C.Structure):
    _fields_ = [('leftLeg', PdLegIn), ('rightLeg', PdLegIn), ('telemetry', C.c_double * 9)]


assert C.sizeof(PdIn) == 952


def make_pd(torque=None, pTarget=None, dTarget=None, pGain=None, dGain=None):
    """10-vectors (left 5, right 5) -> pd_in_t (motorPd branch only)."""
    u = PdIn()
    for name, val in (('torque', torque), ('pTarget', pTarget), ('dTarget', dTarget), ('pGain', pGain), ('dGain', dGain)):
        if val is None:
            continue
        for i in range(5):
            getattr(u.leftLeg.motorPd, name)[i] = float(val[i])
            getattr(u.rightLeg.motorPd, name)[i] = float(val[5 + i])
    return u


def build(ref=False):
    tgt = 'ref' if ref else 'all'
    env = dict(os.environ)
    subprocess.check_call(['make', '-s', '-C', HERE, tgt], env=env)


def lib_path(ref=False):
    return os.path.join(HERE, '_ref', 'liboracle_ref.so') if ref else os.path.join(HERE, '_build', 'liboracle.so')


def load(ref=False):
    p = lib_path(ref)
    if not os.path.exists(p):
        build(ref)
    L = C.CDLL(p)
    L.osim_new.restype = C.c_void_p
    L.osim_new.argtypes = [C.c_char_p]
    L.osim_free.argtypes = [C.c_void_p]
    L.osim_step_pd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.osim_step_pd_no2khz.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.osim_array.restype = C.POINTER(C.c_double)
    L.osim_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
    L.osim_int.argtypes = [C.c_void_p, C.c_char_p]
    L.osim_forward.argtypes = [C.c_void_p]
    L.osim_mj_step.argtypes = [C.c_void_p]
    L.osim_contact.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.osim_drive_filter.restype = C.POINTER(C.c_int)
    L.osim_drive_filter.argtypes = [C.c_void_p]
    for n in ('osim_joint_filter_x', 'osim_joint_filter_y', 'osim_torque_delay'):
        getattr(L, n).restype = C.POINTER(C.c_double)
        getattr(L, n).argtypes = [C.c_void_p]
    L.osim_cassie_out.restype = C.c_void_p
    L.osim_cassie_out.argtypes = [C.c_void_p]
    L.osim_hfield_data.restype = C.POINTER(C.c_float)
    L.osim_hfield_data.argtypes = [C.c_void_p]
    L.osim_run.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int]
    for n in ('osim_foot_forces', 'osim_foot_positions', 'osim_foot_velocities', 'osim_cm_position', 'osim_cm_velocity', 'osim_angular_momentum'):
        getattr(L, n).argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        getattr(L, n).restype = None
    L.osim_heeltoe_forces.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.osim_heeltoe_forces.restype = None
    L.osim_check_obstacle_collision.argtypes = [C.c_void_p]
    L.osim_check_self_collision.argtypes = [C.c_void_p]
    L.osim_geom_collision.argtypes = [C.c_void_p, C.c_int]
    L.osim_model_array.restype = C.POINTER(C.c_double)
    L.osim_model_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
    L.osim_set_const.argtypes = [C.c_void_p]
    L.osim_just_set_const.argtypes = [C.c_void_p]
    L.o_pd_input_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    L.o_core_sim_step.argtypes = [C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_double)]
    return L


class OracleSim:
    """One Cassie environment stepped by the fp64 oracle."""

    def __init__(self, model_path, ref=False):
        self.L = load(ref)
        self.h = self.L.osim_new(model_path.encode())
        if not self.h:
            raise RuntimeError('oracle could not load ' + model_path)
        self.nq, self.nv = self.L.osim_int(self.h, b'nq'), self.L.osim_int(self.h, b'nv')

    def close(self):
        if self.h:
            self.L.osim_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def arr(self, key):
        """live numpy VIEW of an oracle array (writes go through)."""
        n = C.c_int()
        p = self.L.osim_array(self.h, key.encode(), C.byref(n))
        if not p or n.value == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def get_int(self, key):
        return self.L.osim_int(self.h, key.encode())

    def step_pd(self, u, y=None, cassie_out=None):
        self.L.osim_step_pd(self.h, C.byref(u), C.byref(y) if y is not None else None,
                            C.byref(cassie_out) if cassie_out is not None else None)

    def step_pd_no2khz(self, u, y=None):
        self.L.osim_step_pd_no2khz(self.h, C.byref(u), C.byref(y) if y is not None else None)

    def forward(self):
        self.L.osim_forward(self.h)

    def set_caps(self, max_contacts=12, max_rows=48):
        """apply the product's capacity limits so that overflow situations can be compared; (0, 0) = unlimited."""
        self.L.osim_set_caps.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self.L.osim_set_caps(self.h, max_contacts, max_rows)

    def mj_step(self):
        self.L.osim_mj_step(self.h)

    def contacts(self):
        out = []
        buf = (C.c_double * 13)()
        g = (C.c_int * 3)()
        i = 0
        while self.L.osim_contact(self.h, i, buf, g):
            out.append(dict(pos=np.array(buf[0:3]), frame=np.array(buf[3:12]).reshape(3, 3), dist=buf[12], geom1=g[0], geom2=g[1], dim=g[2]))
            i += 1
        return out

    # ---- derived-quantity queries, named like the reference functions they restate (src/cassiemujoco.c:1586-1961)
    def _vec(self, fn, n):
        buf = (C.c_double * n)()
        getattr(self.L, fn)(self.h, buf)
        return np.array(buf[:])

    def foot_forces(self):
        return self._vec('osim_foot_forces', 12)

    def heeltoe_forces(self):
        t, h = (C.c_double * 6)(), (C.c_double * 6)()
        self.L.osim_heeltoe_forces(self.h, t, h)
        return np.array(t[:]), np.array(h[:])

    def foot_positions(self):
        return self._vec('osim_foot_positions', 6)

    def foot_velocities(self):
        return self._vec('osim_foot_velocities', 12)

    def cm_position(self):
        return self._vec('osim_cm_position', 3)

    def cm_velocity(self):
        return self._vec('osim_cm_velocity', 3)

    def angular_momentum(self):
        return self._vec('osim_angular_momentum', 3)

    def check_obstacle_collision(self):
        return bool(self.L.osim_check_obstacle_collision(self.h))

    def check_self_collision(self):
        return bool(self.L.osim_check_self_collision(self.h))

    def geom_collision(self, group):
        return bool(self.L.osim_geom_collision(self.h, group))

    # ---- model constants (the reference hands out c->m->... pointers, src/cassiemujoco.c:1303-1321) and mj_setConst
    def model_arr(self, key):
        """live numpy VIEW of a model array: body_mass, body_ipos, dof_damping, geom_friction, body_invweight0, dof_invweight0, meaninertia."""
        n = C.c_int()
        p = self.L.osim_model_array(self.h, key.encode(), C.byref(n))
        if not p or n.value == 0:
            raise KeyError(key)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def set_const(self):
        self.L.osim_set_const(self.h)

    def just_set_const(self):
        self.L.osim_just_set_const(self.h)

    def efc_J(self):
        n, nv, mv = self.get_int('nefc'), self.nv, self.get_int('MAXV')
        return self.arr('efc_J').reshape(n, mv)[:, :nv].copy() if n else np.zeros((0, nv))
