"""ctypes wrapper of the TEST-ONLY host emulation of the product stepper (tests/emu/emu.cpp)."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(REPO, 'cassie-mujoco-sim_b200', 'csrc')
LIB = os.path.join(HERE, 'emu', '_build', 'libcassie_emu.so')

# debug-dump offsets (devmodel.h)
D = dict(XPOS=0, XQUAT=96, CDOF=224, QM=416, QLD=736, BIAS=1056, PASSIVE=1088, SMOOTH=1120, QACCS=1152, QACC=1184, QFRCC=1216,
         COUNTS=1248, EFC_B=1252, EFC_F=1316, EFC_R=1380, EFC_AREF=1444, SENS=1508, J=1540, SIZE=3600)


def build(force=False):
    srcs = [os.path.join(HERE, 'emu', 'emu.cpp'), os.path.join(CSRC, 'mjcf.cpp'), os.path.join(CSRC, 'step_core.inl'),
            os.path.join(CSRC, 'devmodel.h'), os.path.join(CSRC, 'devbuild.h'), os.path.join(CSRC, 'model.h'), os.path.join(CSRC, 'estimator_host.h'), os.path.join(CSRC, 'cassie_tree_gen.inc')]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(['g++', '-O1', '-std=c++17', '-fPIC', '-shared', '-o', LIB, srcs[0], srcs[1]])


def load():
    build()
    L = C.CDLL(LIB)
    L.emu_new.restype = C.c_void_p
    L.emu_new.argtypes = [C.c_char_p, C.c_int]
    L.emu_free.argtypes = [C.c_void_p]
    L.emu_step.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    L.emu_forward.argtypes = [C.c_void_p]
    L.emu_query.argtypes = [C.c_void_p]
    L.emu_set_const.argtypes = [C.c_void_p]
    L.emu_model_set.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
    L.emu_enable_cenv.argtypes = [C.c_void_p]
    L.emu_plain.argtypes = [C.c_void_p]
    L.emu_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
    L.emu_set.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
    return L


class EmuSim:
    def __init__(self, model_path, fp32=False):
        self.L = load()
        self.h = self.L.emu_new(model_path.encode(), 1 if fp32 else 0)
        if not self.h:
            raise RuntimeError('emu could not load ' + model_path)

    def close(self):
        if self.h:
            self.L.emu_free(self.h)
            self.h = None

    def get(self, name, n=4096):
        buf = np.zeros(n)
        k = self.L.emu_get(self.h, name.encode(), buf.ctypes.data_as(C.POINTER(C.c_double)), n)
        assert k >= 0, name
        return buf[:k]

    def set(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert self.L.emu_set(self.h, name.encode(), a.ctypes.data_as(C.POINTER(C.c_double)), a.size) >= 0

    def step(self, pd50, nticks=1):
        a = np.ascontiguousarray(pd50, dtype=np.float64)
        assert a.size == 50
        self.L.emu_step(self.h, a.ctypes.data_as(C.POINTER(C.c_double)), nticks)

    def forward(self):
        self.L.emu_forward(self.h)

    def query(self):
        self.L.emu_query(self.h)

    def plain(self):
        """switch to the plain kernel instance (no derived-quantity rows, no per-env constants), the one the throughput path runs."""
        self.L.emu_plain(self.h)

    def set_task(self, rows60):
        """taskPd rows (left then right leg: torque, pTarget, dTarget, pGain, dGain [6] each) or None; runs on the extended instance."""
        self.L.emu_set_task.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        if rows60 is None:
            self.L.emu_set_task(self.h, None)
        else:
            a = np.ascontiguousarray(rows60, dtype=np.float64)
            assert a.size == 60
            self.L.emu_set_task(self.h, a.ctypes.data_as(C.POINTER(C.c_double)))

    def set_gait(self, amp, phase, freq):
        """open-loop gait on the motor-PD targets (amp[10], phase[10], freq in Hz) or amp=None: off"""
        self.L.emu_set_gait.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        if amp is None:
            self.L.emu_set_gait(self.h, None)
        else:
            a = np.ascontiguousarray(np.concatenate([amp, phase, [freq]]), dtype=np.float64)
            self.L.emu_set_gait(self.h, a.ctypes.data_as(C.POINTER(C.c_double)))

    def set_geom(self, name, pos=None, quat=None, size=None):
        """move / turn / resize a named geom (cassie_sim_set_geom_name_pos / quat / size) and rebuild the constant block"""
        dp = C.POINTER(C.c_double)
        self.L.emu_set_geom.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
        self.L.emu_geom_id.argtypes = [C.c_void_p, C.c_char_p]
        g = self.L.emu_geom_id(self.h, name.encode())
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (pos, quat, size)]
        assert self.L.emu_set_geom(self.h, g, *[None if a is None else a.ctypes.data_as(dp) for a in arrs]) == 0, name
        return g

    def enable_est(self, on=True):
        """in-kernel estimator (forces + filters) of the extended instance; enabling restarts it"""
        self.L.emu_enable_est.argtypes = [C.c_void_p, C.c_int]
        self.L.emu_enable_est(self.h, 1 if on else 0)

    def enable_cenv(self):
        self.L.emu_enable_cenv(self.h)

    def set_const(self):
        self.L.emu_set_const(self.h)

    def model_set(self, what, values):
        a = np.ascontiguousarray(values, dtype=np.float64)
        assert self.L.emu_model_set(self.h, what.encode(), a.ctypes.data_as(C.POINTER(C.c_double)), a.size) == 0, what


def _emu_set_hfield(self, data):
    a = np.ascontiguousarray(data, dtype=np.float32)
    self.L.emu_set_hfield.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
    self.L.emu_set_hfield(self.h, a.ctypes.data_as(C.POINTER(C.c_float)), a.size)


EmuSim.set_hfield = _emu_set_hfield
