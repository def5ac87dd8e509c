"""Study aid for the reference's closed estimator (test infrastructure; needs oracle/_ref/libprobe_est.so = `make -C oracle probe`)."""
import ctypes as C
import os
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
_L = None


def lib():
    global _L
    if _L is None:
        _L = C.CDLL(os.path.join(HERE, '_ref', 'libprobe_est.so'))
        _L.probe_est.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        _L.probe_est_seq.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
    return _L


def est(inp, n=1):
    a = np.ascontiguousarray(inp, dtype=np.float64)
    out = np.zeros(123)
    lib().probe_est(a.ctypes.data_as(C.POINTER(C.c_double)), n, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:105]


def est_seq(inps):
    a = np.ascontiguousarray(inps, dtype=np.float64)
    T = a.shape[0]
    out = np.zeros((T, 105))
    lib().probe_est_seq(a.ctypes.data_as(C.POINTER(C.c_double)), T, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def nominal():
    x = np.zeros(45)
    x[32] = 1
    x[:10] = [0.0045, 0, 0.4973, -1.1997, -1.5968, -0.0045, 0, 0.4973, -1.1997, -1.5968]
    x[20:26] = [0, 1.4267, -1.5968, 0, 1.4267, -1.5968]
    return x


def Rz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.]])


def frame(xa, ya):
    xa, ya = np.array(xa, float), np.array(ya, float)
    return np.stack([xa, ya, np.cross(xa, ya)], axis=1)


def fk(side, m, shin, tars):
    """foot point and foot frame in the pelvis frame; m = hipRoll, hipYaw, hipPitch, knee, foot (motor positions)"""
    sgn = 1 if side == 0 else -1
    steps = [((0.021, 0.135 * sgn, 0), frame((0, 0, -1), (0, 1, 0)), m[0]), ((0, 0, -0.07), frame((0, 0, 1), (0, 1, 0)), m[1]),
             ((0, 0, -0.09), frame((0, 0, -1), (1, 0, 0)), m[2]), ((0.12, 0, 0.0045 * sgn), np.eye(3), m[3]),
             ((0.06068, 0.04741, 0), np.eye(3), shin), ((0.43476, 0.02, 0), np.eye(3), tars), ((0.408, -0.04, 0), np.eye(3), m[4])]
    R, p = np.eye(3), np.zeros(3)
    for pos, F, q in steps:
        p = p + R @ np.array(pos)
        R = R @ F @ Rz(q)
    c40, s40 = np.cos(np.deg2rad(40)), np.sin(np.deg2rad(40))
    Roff = np.array([[-c40, 0, -s40], [s40, 0, -c40], [0, -1, 0.]])
    return p + R @ np.array([0.01762, 0.05219, 0]), R @ Roff


def est_mem(inps):
    """(outputs [T,105], memory snapshots [T+1, nbytes] of the estimator block) for a sequence of inputs"""
    a = np.ascontiguousarray(inps, dtype=np.float64)
    T = a.shape[0]
    L = lib()
    L.probe_est_mem.restype = C.c_long
    L.probe_est_mem.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_void_p, C.c_long, C.POINTER(C.c_double)]
    sz = L.probe_est_mem(a.ctypes.data_as(C.POINTER(C.c_double)), 0, None, 0, None)
    mem = np.zeros((T + 1, sz), dtype=np.uint8)
    out = np.zeros((T, 105))
    L.probe_est_mem(a.ctypes.data_as(C.POINTER(C.c_double)), T, mem.ctypes.data, sz, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out, mem
