"""cassie-mujoco-sim_b200 -- Python host side of the B200-native batched Cassie stepper.

Mirrors the reference's Python surface for the one path this repo accelerates
(/root/reference/example/cassiemujoco.py:31-173: CassieSim.step_pd / qpos / qvel / set_qpos / apply_force ...)
on top of the C-ABI in include/cassie_b200.h, and adds CassieBatch for the batched entry points.
There is no CPU fallback: loading fails loudly if the CUDA library is missing, and cassie_batch_init fails if no GPU is visible.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcassie_b200.so')
MODEL_DIR = os.path.join(_HERE, 'models')
FP32, FP64 = 0, 1
PD_WIDTH, OBS_WIDTH, AUX_WIDTH, EST_WIDTH = 52, 112, 64, 16
# slices of a derived-quantity row (CASSIE_AUX_* in include/cassie_b200.h)
AUX = dict(foot_force=slice(0, 12), toe_force=slice(12, 18), heel_force=slice(18, 24), foot_pos=slice(24, 30), foot_vel=slice(30, 42),
           cm_pos=slice(42, 45), cm_vel=slice(45, 48), angmom=slice(48, 51), obstacle=51, self_collision=52, group_mask=53, ncon=54)
# observation row layout (cassie_batch_get_obs)
OBS = dict(motor_pos=slice(0, 10), motor_vel=slice(10, 20), motor_torque=slice(20, 30), joint_pos=slice(30, 36), joint_vel=slice(36, 42),
           quat=slice(42, 46), gyro=slice(46, 49), accel=slice(49, 52), mag=slice(52, 55), time=55,
           # decoded estimator (state_output_step): stateless part ...
           est_accel=slice(56, 59), left_foot=slice(60, 73), right_foot=slice(73, 86), est_quat=slice(86, 90),
           # ... and, once the in-kernel estimator is on, the force model and the filters
           est_position=slice(96, 99), est_velocity=slice(99, 102), est_external_force=slice(102, 105), est_terrain_height=105,
           est_left_toe_force=slice(106, 109), est_right_toe_force=slice(109, 112))


# ---------------------------------------------------------------- ctypes mirrors of the bus structs (include/cassie_bus.h)
class pd_motor_in_t(C.Structure):
    _fields_ = [(n, C.c_double * 5) for n in ('torque', 'pTarget', 'dTarget', 'pGain', 'dGain')]


class pd_task_in_t(C.Structure):
    _fields_ = [(n, C.c_double * 6) for n in ('torque', 'pTarget', 'dTarget', 'pGain', 'dGain')]


class pd_leg_in_t(C.Structure):
    _fields_ = [('taskPd', pd_task_in_t), ('motorPd', pd_motor_in_t)]


class pd_in_t(C.Structure):
    _fields_ = [('leftLeg', pd_leg_in_t), ('rightLeg', pd_leg_in_t), ('telemetry', C.c_double * 9)]


class state_battery_out_t(C.Structure):
    _fields_ = [('stateOfCharge', C.c_double), ('current', C.c_double)]


class state_foot_out_t(C.Structure):
    _fields_ = [('position', C.c_double * 3), ('orientation', C.c_double * 4), ('footRotationalVelocity', C.c_double * 3),
                ('footTranslationalVelocity', C.c_double * 3), ('toeForce', C.c_double * 3), ('heelForce', C.c_double * 3)]


class state_joint_out_t(C.Structure):
    _fields_ = [('position', C.c_double * 6), ('velocity', C.c_double * 6)]


class state_motor_out_t(C.Structure):
    _fields_ = [('position', C.c_double * 10), ('velocity', C.c_double * 10), ('torque', C.c_double * 10)]


class state_pelvis_out_t(C.Structure):
    _fields_ = [('position', C.c_double * 3), ('orientation', C.c_double * 4), ('rotationalVelocity', C.c_double * 3),
                ('translationalVelocity', C.c_double * 3), ('translationalAcceleration', C.c_double * 3),
                ('externalMoment', C.c_double * 3), ('externalForce', C.c_double * 3)]


class state_radio_out_t(C.Structure):
    _fields_ = [('channel', C.c_double * 16), ('signalGood', C.c_bool)]


class state_terrain_out_t(C.Structure):
    _fields_ = [('height', C.c_double), ('slope', C.c_double * 2)]


class state_out_t(C.Structure):
    _fields_ = [('pelvis', state_pelvis_out_t), ('leftFoot', state_foot_out_t), ('rightFoot', state_foot_out_t),
                ('terrain', state_terrain_out_t), ('motor', state_motor_out_t), ('joint', state_joint_out_t),
                ('radio', state_radio_out_t), ('battery', state_battery_out_t)]


assert C.sizeof(pd_in_t) == 952 and C.sizeof(state_out_t) == 992

_lib = None


def build(force=False, verbose=False):
    """compile libcassie_b200.so in-tree (nvcc, sm_100a)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('_cassie_b200_build', os.path.join(_HERE, 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


def lib():
    """the loaded C-ABI library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('libcassie_b200.so is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). '
                           'The batched stepper has no CPU fallback.')
    L = C.CDLL(LIB_PATH)
    vp, ci, cd = C.c_void_p, C.c_int, C.POINTER(C.c_double)
    L.cassie_b200_last_error.restype = C.c_char_p
    L.cassie_batch_init.restype = vp
    L.cassie_batch_init.argtypes = [C.c_char_p, ci, ci, ci]
    L.cassie_batch_free.argtypes = [vp]
    for n in ('cassie_batch_nenv', 'cassie_batch_nq', 'cassie_batch_nv', 'cassie_batch_precision'):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = ci
    L.cassie_batch_launch_count.argtypes = [vp]
    L.cassie_batch_launch_count.restype = C.c_long
    L.cassie_batch_reset.argtypes = [vp, C.c_void_p]
    L.cassie_sim_step_pd_batch.argtypes = [vp, C.c_void_p, C.c_void_p]
    L.cassie_batch_set_pd.argtypes = [vp, cd]
    L.cassie_batch_step.argtypes = [vp, ci]
    for n in ('cassie_batch_sync', 'cassie_batch_forward', 'cassie_batch_clear_forces', 'cassie_batch_integrate_pos'):
        getattr(L, n).argtypes = [vp]
    for n in ('cassie_batch_get_qpos', 'cassie_batch_set_qpos', 'cassie_batch_get_qvel', 'cassie_batch_set_qvel', 'cassie_batch_get_time',
              'cassie_batch_get_obs'):
        getattr(L, n).argtypes = [vp, cd]
    L.cassie_batch_apply_force.argtypes = [vp, cd, C.c_char_p]
    L.cassie_batch_apply_force.restype = ci
    L.cassie_batch_set_hfielddata.argtypes = [vp, C.POINTER(C.c_float), ci]
    L.cassie_batch_set_hfielddata.restype = ci
    for n in ('cassie_batch_hfield_nrow', 'cassie_batch_hfield_ncol'):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = ci
    L.cassie_batch_device_ptr.argtypes = [vp, C.c_char_p]
    L.cassie_batch_device_ptr.restype = vp
    L.cassie_batch_set_stream.argtypes = [vp, vp]
    L.cassie_batch_get_stream.argtypes = [vp]
    L.cassie_batch_get_stream.restype = vp
    L.cassie_batch_get_counters.argtypes = [vp, C.POINTER(ci)]
    L.cassie_batch_debug_dump.argtypes = [vp, ci, cd, ci]
    L.cassie_batch_debug_dump.restype = ci
    # legacy verbs
    L.cassie_mujoco_init.argtypes = [C.c_char_p]
    L.cassie_mujoco_init.restype = C.c_bool
    L.cassie_sim_init.argtypes = [C.c_char_p, C.c_bool]
    L.cassie_sim_init.restype = vp
    L.cassie_sim_free.argtypes = [vp]
    L.cassie_sim_step_pd.argtypes = [vp, C.POINTER(state_out_t), C.POINTER(pd_in_t)]
    for n in ('cassie_sim_time', 'cassie_sim_qpos', 'cassie_sim_qvel', 'cassie_sim_timestep', 'cassie_state_time', 'cassie_state_qpos', 'cassie_state_qvel'):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = cd
    L.cassie_sim_step.argtypes = [vp, vp, vp]
    L.cassie_sim_step_pd_no2khz.argtypes = [vp, C.POINTER(state_out_t), C.POINTER(pd_in_t)]
    L.cassie_sim_set_timestep.argtypes = [vp, C.c_double]
    L.cassie_sim_forward.argtypes = [vp]
    L.cassie_sim_forward.restype = ci
    for n in ('cassie_sim_hold', 'cassie_sim_release', 'cassie_state_free'):
        getattr(L, n).argtypes = [vp]
    for n in ('cassie_get_state', 'cassie_set_state', 'cassie_sim_copy', 'cassie_state_copy'):
        getattr(L, n).argtypes = [vp, vp]
    L.cassie_state_alloc.restype = vp
    L.cassie_sim_duplicate.argtypes = [vp]
    L.cassie_sim_duplicate.restype = vp
    for n in ('cassie_sim_nv', 'cassie_sim_nq'):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = ci
    L.cassie_sim_apply_force.argtypes = [vp, cd, C.c_char_p]
    L.cassie_sim_clear_forces.argtypes = [vp]
    L.cassie_sim_full_reset.argtypes = [vp]
    L.cassie_sim_radio.argtypes = [vp, cd]
    for n in ('cassie_batch_enable_aux',):
        getattr(L, n).argtypes = [vp, ci]
        getattr(L, n).restype = ci
    L.cassie_batch_enable_estimator_forces.argtypes = [vp, ci]
    L.cassie_batch_enable_estimator_forces.restype = ci
    L.cassie_batch_enable_estimator_filter.argtypes = [vp, ci]
    L.cassie_batch_enable_estimator_filter.restype = ci
    L.cassie_batch_reset_estimator.argtypes = [vp, C.c_void_p]
    L.cassie_batch_reset_estimator.restype = ci
    L.cassie_batch_enable_estimator_device.argtypes = [vp, ci]
    L.cassie_batch_enable_estimator_device.restype = ci
    L.cassie_batch_get_estimator.argtypes = [vp, cd]
    L.cassie_batch_get_estimator.restype = ci
    L.cassie_batch_set_pd_gait.argtypes = [vp, cd, cd, cd]
    L.cassie_batch_set_pd_gait.restype = ci
    L.cassie_batch_set_task_pd.argtypes = [vp, cd]
    L.cassie_batch_set_task_pd.restype = ci
    L.cassie_batch_get_aux.argtypes = [vp, cd]
    L.cassie_batch_get_aux.restype = ci
    L.cassie_batch_query.argtypes = [vp]
    L.cassie_batch_query.restype = ci
    L.cassie_batch_row_width.argtypes = [vp, C.c_char_p]
    L.cassie_batch_row_width.restype = ci
    for n in ('cassie_batch_nbody', 'cassie_batch_ngeom'):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = ci
    for what in ('body_mass', 'body_ipos', 'dof_damping', 'geom_friction'):
        for op in ('set', 'get'):
            f = getattr(L, 'cassie_batch_%s_%s' % (op, what))
            f.argtypes = [vp, cd]
            f.restype = ci
    L.cassie_batch_set_const.argtypes = [vp, C.c_void_p, ci]
    L.cassie_batch_set_const.restype = ci
    L.cassie_sim_params.argtypes = [vp, C.POINTER(ci)]
    for n in ('cassie_sim_dof_damping', 'cassie_sim_body_mass', 'cassie_sim_body_ipos', 'cassie_sim_geom_friction'):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = cd
    for n in ('cassie_sim_set_dof_damping', 'cassie_sim_set_body_mass', 'cassie_sim_set_body_ipos', 'cassie_sim_set_geom_friction'):
        getattr(L, n).argtypes = [vp, cd]
        getattr(L, n).restype = None
    for n in ('cassie_sim_set_dof_name_damping', 'cassie_sim_set_body_name_ipos', 'cassie_sim_set_geom_name_friction'):
        getattr(L, n).argtypes = [vp, C.c_char_p, cd]
        getattr(L, n).restype = None
    for n in ('cassie_sim_get_dof_name_damping', 'cassie_sim_get_body_name_ipos', 'cassie_sim_get_geom_name_friction'):
        getattr(L, n).argtypes = [vp, C.c_char_p]
        getattr(L, n).restype = cd
    L.cassie_sim_get_joint_num_dof.argtypes = [vp, C.c_char_p]
    L.cassie_sim_get_joint_num_dof.restype = ci
    L.cassie_sim_set_body_name_mass.argtypes = [vp, C.c_char_p, C.c_double]
    L.cassie_sim_set_body_name_mass.restype = None
    L.cassie_sim_get_body_name_mass.argtypes = [vp, C.c_char_p]
    L.cassie_sim_get_body_name_mass.restype = C.c_double
    for n in ('cassie_sim_set_const', 'cassie_sim_just_set_const'):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = None
    for n in ('cassie_sim_foot_forces', 'cassie_sim_foot_positions', 'cassie_sim_foot_velocities', 'cassie_sim_cm_position', 'cassie_sim_cm_velocity',
              'cassie_sim_angular_momentum'):
        getattr(L, n).argtypes = [vp, cd]
        getattr(L, n).restype = None
    L.cassie_sim_heeltoe_forces.argtypes = [vp, cd, cd]
    L.cassie_sim_heeltoe_forces.restype = None
    for n in ('cassie_sim_check_obstacle_collision', 'cassie_sim_check_self_collision'):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = C.c_bool
    L.cassie_sim_geom_collision.argtypes = [vp, ci]
    L.cassie_sim_geom_collision.restype = C.c_bool
    _lib = L
    return L


def model_path(name='cassie'):
    """compiled model table shipped with the package (used where the reference's MJCF checkout is absent)."""
    return os.path.join(MODEL_DIR, name + '.cmodel')


def pd_rows(n, torque=None, pTarget=None, dTarget=None, pGain=None, dGain=None):
    """[n][52] compact motor-PD rows from 10-vectors (broadcast over envs) or [n][10] arrays."""
    rows = np.zeros((n, PD_WIDTH))
    for k, v in enumerate((torque, pTarget, dTarget, pGain, dGain)):
        if v is not None:
            rows[:, 10 * k:10 * k + 10] = np.asarray(v, dtype=np.float64)
    return rows


def _last_error():
    return lib().cassie_b200_last_error().decode()


class CassieBatch:
    """n_env independent Cassie simulators advanced in lock-step on one GPU."""

    def __init__(self, n_env, modelfile=None, device=0, precision=FP32):
        self.L = lib()
        self.n = int(n_env)
        path = modelfile or model_path()
        self.h = self.L.cassie_batch_init(path.encode(), self.n, int(device), int(precision))
        if not self.h:
            raise RuntimeError('cassie_batch_init failed: ' + _last_error())
        self.nq, self.nv = self.L.cassie_batch_nq(self.h), self.L.cassie_batch_nv(self.h)
        self.precision = precision

    def close(self):
        if getattr(self, 'h', None):
            self.L.cassie_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _dp(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    # ---- stepping
    def set_pd(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        assert rows.shape == (self.n, PD_WIDTH)
        self.L.cassie_batch_set_pd(self.h, self._dp(rows))

    def set_task_pd(self, rows):
        """taskPd rows [n, 60] (per leg torque, pTarget, dTarget, pGain, dGain [6]) or None to switch the branch off."""
        if rows is None:
            rc = self.L.cassie_batch_set_task_pd(self.h, None)
        else:
            a = np.ascontiguousarray(rows, dtype=np.float64).reshape(self.n, 60)
            rc = self.L.cassie_batch_set_task_pd(self.h, self._dp(a))
        if rc != 0:
            raise RuntimeError(_last_error())

    def set_pd_gait(self, amp=None, freq=None, phase=None):
        """open-loop gait on the motor-PD targets: pTarget_i(t) = row pTarget_i + amp[e, i] sin(2 pi freq[e] t + phase[e, i]), evaluated inside the kernel
        every control tick (t = ticks since reset / since this call x 0.5 ms); amp=None switches it off."""
        if amp is None:
            rc = self.L.cassie_batch_set_pd_gait(self.h, None, None, None)
        else:
            a = np.ascontiguousarray(np.broadcast_to(amp, (self.n, 10)), dtype=np.float64)
            f = np.ascontiguousarray(np.broadcast_to(freq, (self.n,)), dtype=np.float64)
            p = np.ascontiguousarray(np.broadcast_to(phase, (self.n, 10)), dtype=np.float64)
            rc = self.L.cassie_batch_set_pd_gait(self.h, self._dp(a), self._dp(f), self._dp(p))
        if rc != 0:
            raise RuntimeError(_last_error())

    def step(self, nticks=1):
        self.L.cassie_batch_step(self.h, int(nticks))

    def sync(self):
        self.L.cassie_batch_sync(self.h)

    def step_pd(self, pd_in_array, want_state=True):
        """AoS compatibility path: (pd_in_t * n) in, (state_out_t * n) out (cassie_sim_step_pd_batch)."""
        out = (state_out_t * self.n)() if want_state else None
        self.L.cassie_sim_step_pd_batch(self.h, C.byref(pd_in_array), C.byref(out) if want_state else None)
        return out

    def reset(self, mask=None):
        if mask is None:
            self.L.cassie_batch_reset(self.h, None)
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            self.L.cassie_batch_reset(self.h, m.ctypes.data_as(C.c_void_p))

    def forward(self):
        self.L.cassie_batch_forward(self.h)

    def integrate_pos(self):
        self.L.cassie_batch_integrate_pos(self.h)

    # ---- state
    def _get(self, fn, w):
        out = np.zeros((self.n, w))
        fn(self.h, self._dp(out))
        return out

    def qpos(self):
        return self._get(self.L.cassie_batch_get_qpos, self.nq)

    def qvel(self):
        return self._get(self.L.cassie_batch_get_qvel, self.nv)

    def time(self):
        return self._get(self.L.cassie_batch_get_time, 1)[:, 0]

    def obs(self):
        return self._get(self.L.cassie_batch_get_obs, OBS_WIDTH)

    # ---- the estimator's host-side part for step_pd(): toe / heel forces, and the filters behind pelvis.position / translationalVelocity /
    # externalForce and terrain.height (one filter per environment, advanced once per step_pd call as the reference's 2 kHz estimator is)
    def enable_estimator(self, forces=True, filters=True):
        """HOST-side checker of the estimator (one filter object per environment, advanced per step_pd call); the product path is the
        in-kernel estimator (enable_estimator_device), which step_pd switches on by itself"""
        self.L.cassie_batch_enable_estimator_forces(self.h, 1 if (forces or filters) else 0)
        self.L.cassie_batch_enable_estimator_filter(self.h, 1 if filters else 0)

    def enable_estimator_device(self, on=True):
        """the estimator inside the step kernel: columns 96..111 of obs() / estimator() follow every tick of every launch; step_pd switches it
        on by itself at its first call, enable_estimator_device(False) keeps it off"""
        if self.L.cassie_batch_enable_estimator_device(self.h, 1 if on else 0) != 0:
            raise RuntimeError(_last_error())

    def estimator(self):
        """[n, 16]: position 3, translationalVelocity 3, externalForce 3, terrain.height, toeForce L 3, R 3"""
        out = np.zeros((self.n, EST_WIDTH))
        if self.L.cassie_batch_get_estimator(self.h, self._dp(out)) != 0:
            raise RuntimeError(_last_error())
        return out

    def reset_estimator(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.cassie_batch_reset_estimator(self.h, None if m is None else m.ctypes.data)

    # ---- derived quantities (reference: the read-only queries of example/cassiemujoco.py:214-306, 815-819), one row per environment
    def enable_aux(self, on=True):
        if self.L.cassie_batch_enable_aux(self.h, 1 if on else 0) != 0:
            raise RuntimeError(_last_error())

    def aux(self):
        """[n, AUX_WIDTH] rows; slices are named in AUX (e.g. rows[:, AUX['foot_force']])."""
        out = np.zeros((self.n, AUX_WIDTH))
        if self.L.cassie_batch_get_aux(self.h, self._dp(out)) != 0:
            raise RuntimeError(_last_error())
        return out

    # ---- per-environment model constants (domain randomisation; reference: example/cassiemujoco.py:517-610 on one env)
    def _model_width(self, what):
        nb, ng = self.L.cassie_batch_nbody(self.h), self.L.cassie_batch_ngeom(self.h)
        return dict(body_mass=nb, body_ipos=3 * nb, dof_damping=self.nv, geom_friction=3 * ng)[what]

    def set_model(self, what, rows):
        """what in body_mass [n, nbody], body_ipos [n, 3 nbody], dof_damping [n, nv], geom_friction [n, 3 ngeom] (reference numbering)."""
        a = np.ascontiguousarray(rows, dtype=np.float64).reshape(self.n, self._model_width(what))
        if getattr(self.L, 'cassie_batch_set_' + what)(self.h, self._dp(a)) != 0:
            raise RuntimeError(_last_error())

    def get_model(self, what):
        out = np.zeros((self.n, self._model_width(what)))
        if getattr(self.L, 'cassie_batch_get_' + what)(self.h, self._dp(out)) != 0:
            raise RuntimeError(_last_error())
        return out

    def set_const(self, mask=None, reset_state=False):
        """mj_setConst on the device for the masked environments (all if None); reset_state=True adds cassie_sim_set_const's state reset."""
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        if self.L.cassie_batch_set_const(self.h, None if m is None else m.ctypes.data_as(C.c_void_p), 1 if reset_state else 0) != 0:
            raise RuntimeError(_last_error())

    def query(self):
        """refresh the centre-of-mass slots of the aux rows for the CURRENT state (nothing else is written)."""
        if self.L.cassie_batch_query(self.h) != 0:
            raise RuntimeError(_last_error())

    def set_qpos(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == (self.n, self.nq)
        self.L.cassie_batch_set_qpos(self.h, self._dp(q))

    def set_qvel(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.shape == (self.n, self.nv)
        self.L.cassie_batch_set_qvel(self.h, self._dp(v))

    def apply_force(self, xfrc, body_name='cassie-pelvis'):
        x = np.ascontiguousarray(np.broadcast_to(np.asarray(xfrc, dtype=np.float64), (self.n, 6)))
        return self.L.cassie_batch_apply_force(self.h, self._dp(x), body_name.encode())

    def clear_forces(self):
        self.L.cassie_batch_clear_forces(self.h)

    def set_hfield_data(self, data):
        """data: [K, nrow, ncol] (or [nrow, ncol]) normalised elevations; env e uses terrain e % K."""
        a = np.ascontiguousarray(data, dtype=np.float32)
        nrow, ncol = self.L.cassie_batch_hfield_nrow(self.h), self.L.cassie_batch_hfield_ncol(self.h)
        a = a.reshape(-1, nrow, ncol)
        if self.L.cassie_batch_set_hfielddata(self.h, a.ctypes.data_as(C.POINTER(C.c_float)), a.shape[0]) != 0:
            raise RuntimeError(_last_error())

    def get_state(self, snap=None):
        """device-resident snapshot of every row array (qpos, qvel, warm start, controller / sensor state, filters, forces, estimator, observation rows)"""
        self.L.cassie_batch_state_alloc.restype = C.c_void_p
        self.L.cassie_batch_state_alloc.argtypes = [C.c_void_p]
        self.L.cassie_batch_get_state.argtypes = [C.c_void_p, C.c_void_p]
        snap = snap or self.L.cassie_batch_state_alloc(self.h)
        if not snap or self.L.cassie_batch_get_state(self.h, C.c_void_p(snap)) != 0:
            raise RuntimeError(_last_error())
        return snap

    def set_state(self, snap, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.cassie_batch_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        if self.L.cassie_batch_set_state(self.h, C.c_void_p(snap), None if m is None else m.ctypes.data) != 0:
            raise RuntimeError(_last_error())

    def free_state(self, snap):
        self.L.cassie_batch_state_free.argtypes = [C.c_void_p, C.c_void_p]
        self.L.cassie_batch_state_free(self.h, C.c_void_p(snap))

    def set_timestep(self, dt):
        self.L.cassie_batch_set_timestep.argtypes = [C.c_void_p, C.c_double]
        if self.L.cassie_batch_set_timestep(self.h, float(dt)) != 0:
            raise RuntimeError(_last_error())

    def row_width(self, field):
        """row width (elements) of a device array: qpos 36 (44 with the extra free body), qvel 32 (40), pd 52, obs 112, xfrc 8, aux 64"""
        self.L.cassie_batch_row_width.argtypes = [C.c_void_p, C.c_char_p]
        return int(self.L.cassie_batch_row_width(self.h, field.encode()))

    def counters(self):
        out = np.zeros((self.n, 8), dtype=np.int32)
        self.L.cassie_batch_get_counters(self.h, out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def debug_dump(self, env=0):
        out = np.zeros(3600)
        k = self.L.cassie_batch_debug_dump(self.h, env, self._dp(out), out.size)
        return out if k > 0 else None

    def launch_count(self):
        return int(self.L.cassie_batch_launch_count(self.h))

    def device_ptr(self, field):
        return self.L.cassie_batch_device_ptr(self.h, field.encode())

    def set_stream(self, cuda_stream_ptr):
        self.L.cassie_batch_set_stream(self.h, C.c_void_p(cuda_stream_ptr))

    def torch_view(self, field):
        """zero-copy torch tensor over a device state array (qpos [n,36], qvel [n,32], pd [n,52], obs [n,112]).

        The arrays are written on the batch's stream.  Unless that stream is torch's current stream (set_stream(torch.cuda.current_stream().cuda_stream)),
        call sync() before reading the view: torch ops on another stream are not ordered after an in-flight step."""
        import torch
        self.L.cassie_batch_row_width.argtypes = [C.c_void_p, C.c_char_p]
        width = self.L.cassie_batch_row_width(self.h, field.encode())
        dt, isz, ts = (np.float32, 4, '<f4') if self.precision == FP32 else (np.float64, 8, '<f8')

        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = dict(shape=(self.n, width), typestr=ts, data=(self.device_ptr(field), False), version=2)
        return torch.as_tensor(a, device='cuda')


class cassie_user_in_t(C.Structure):
    _fields_ = [('torque', C.c_double * 10), ('telemetry', C.c_short * 9)]


class elmo_out_t(C.Structure):
    _fields_ = [('statusWord', C.c_ushort), ('position', C.c_double), ('velocity', C.c_double), ('torque', C.c_double), ('driveTemperature', C.c_double),
                ('dcLinkVoltage', C.c_double), ('torqueLimit', C.c_double), ('gearRatio', C.c_double)]


class cassie_joint_out_t(C.Structure):
    _fields_ = [('position', C.c_double), ('velocity', C.c_double)]


class cassie_leg_out_t(C.Structure):
    _fields_ = [(n, elmo_out_t) for n in ('hipRollDrive', 'hipYawDrive', 'hipPitchDrive', 'kneeDrive', 'footDrive')] + \
               [(n, cassie_joint_out_t) for n in ('shinJoint', 'tarsusJoint', 'footJoint')] + \
               [('medullaCounter', C.c_ubyte), ('medullaCpuLoad', C.c_ushort), ('reedSwitchState', C.c_bool)]


class battery_out_t(C.Structure):
    _fields_ = [('dataGood', C.c_bool), ('stateOfCharge', C.c_double), ('voltage', C.c_double * 12), ('current', C.c_double), ('temperature', C.c_double * 4)]


class radio_out_t(C.Structure):
    _fields_ = [('radioReceiverSignalGood', C.c_bool), ('receiverMedullaSignalGood', C.c_bool), ('channel', C.c_double * 16)]


class target_pc_out_t(C.Structure):
    _fields_ = [('etherCatStatus', C.c_int * 6), ('etherCatNotifications', C.c_int * 21), ('taskExecutionTime', C.c_double), ('overloadCounter', C.c_uint), ('cpuTemperature', C.c_double)]


class vectornav_out_t(C.Structure):
    _fields_ = [('dataGood', C.c_bool), ('vpeStatus', C.c_ushort), ('pressure', C.c_double), ('temperature', C.c_double), ('magneticField', C.c_double * 3),
                ('angularVelocity', C.c_double * 3), ('linearAcceleration', C.c_double * 3), ('orientation', C.c_double * 4)]


class cassie_pelvis_out_t(C.Structure):
    _fields_ = [('targetPc', target_pc_out_t), ('battery', battery_out_t), ('radio', radio_out_t), ('vectorNav', vectornav_out_t), ('medullaCounter', C.c_ubyte),
                ('medullaCpuLoad', C.c_ushort), ('bleederState', C.c_bool), ('leftReedSwitchState', C.c_bool), ('rightReedSwitchState', C.c_bool), ('vtmTemperature', C.c_double)]


class cassie_out_t(C.Structure):
    _fields_ = [('pelvis', cassie_pelvis_out_t), ('leftLeg', cassie_leg_out_t), ('rightLeg', cassie_leg_out_t), ('isCalibrated', C.c_bool), ('messages', C.c_short * 4)]


assert C.sizeof(cassie_out_t) == 1336 and C.sizeof(cassie_user_in_t) == 104


class CassieState:
    """cassie_state_t (reference: example/cassiemujoco.py CassieState): a full dynamic-state snapshot of a CassieSim"""

    def __init__(self):
        self.L = lib()
        self.s = self.L.cassie_state_alloc()

    def time(self):
        return self.L.cassie_state_time(self.s)[0]

    def qpos(self):
        return np.array(self.L.cassie_state_qpos(self.s)[:35])

    def qvel(self):
        return np.array(self.L.cassie_state_qvel(self.s)[:32])

    def set_time(self, t):
        self.L.cassie_state_time(self.s)[0] = t

    def set_qpos(self, q):
        p = self.L.cassie_state_qpos(self.s)
        for i in range(len(q)):
            p[i] = q[i]

    def set_qvel(self, v):
        p = self.L.cassie_state_qvel(self.s)
        for i in range(len(v)):
            p[i] = v[i]

    def close(self):
        if getattr(self, 's', None):
            self.L.cassie_state_free(self.s)
            self.s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CassieSim:
    """Legacy single-environment object (reference: example/cassiemujoco.py:31-173), stepped by the same CUDA kernels."""

    def __init__(self, modelfile=None, reinit=True):   # the reference wrapper always passes reinit=True (example/cassiemujoco.py:44)
        self.L = lib()
        path = modelfile or model_path()
        self.c = self.L.cassie_sim_init(path.encode(), bool(reinit))
        if not self.c:
            raise RuntimeError('cassie_sim_init failed: ' + _last_error())
        self.nq, self.nv = self.L.cassie_sim_nq(self.c), self.L.cassie_sim_nv(self.c)

    def step_pd(self, u):
        y = state_out_t()
        self.L.cassie_sim_step_pd(self.c, C.byref(y), C.byref(u))
        return y

    def time(self):
        return self.L.cassie_sim_time(self.c)[0]

    def qpos(self):
        return np.array(self.L.cassie_sim_qpos(self.c)[:self.nq])

    def qvel(self):
        return np.array(self.L.cassie_sim_qvel(self.c)[:self.nv])

    def set_time(self, t):
        self.L.cassie_sim_time(self.c)[0] = t

    def set_qpos(self, qpos):
        p = self.L.cassie_sim_qpos(self.c)
        for i in range(min(len(qpos), self.nq)):
            p[i] = qpos[i]

    def set_qvel(self, qvel):
        p = self.L.cassie_sim_qvel(self.c)
        for i in range(min(len(qvel), self.nv)):
            p[i] = qvel[i]

    def apply_force(self, xfrc, body_name='cassie-pelvis'):
        a = (C.c_double * 6)(*[float(x) for x in xfrc])
        self.L.cassie_sim_apply_force(self.c, a, body_name.encode())

    def clear_forces(self):
        self.L.cassie_sim_clear_forces(self.c)

    def full_reset(self):
        self.L.cassie_sim_full_reset(self.c)

    # ---- the verbs around the hot path (reference: example/cassiemujoco.py:64-173)
    def step(self, u):
        """torque-level step: u = cassie_user_in_t; returns the cassie_out_t of this tick"""
        y = cassie_out_t()
        self.L.cassie_sim_step(self.c, C.addressof(y), C.addressof(u))
        return y

    def step_pd_no2khz(self, u):
        y = state_out_t()
        self.L.cassie_sim_step_pd_no2khz(self.c, C.byref(y), C.byref(u))
        return y

    def get_cassie_out(self):
        self.L.cassie_sim_get_cassie_out.restype = cassie_out_t
        self.L.cassie_sim_get_cassie_out.argtypes = [C.c_void_p]
        return self.L.cassie_sim_get_cassie_out(self.c)

    def timestep(self):
        return self.L.cassie_sim_timestep(self.c)[0]

    def set_timestep(self, dt):
        self.L.cassie_sim_set_timestep(self.c, float(dt))

    def forward(self):
        return self.L.cassie_sim_forward(self.c)

    def hold(self):
        self.L.cassie_sim_hold(self.c)

    def release(self):
        self.L.cassie_sim_release(self.c)

    def get_state(self, s=None):
        s = s or CassieState()
        self.L.cassie_get_state(self.c, s.s)
        return s

    def set_state(self, s):
        self.L.cassie_set_state(self.c, s.s)

    def duplicate(self):
        d = CassieSim.__new__(CassieSim)
        d.L = self.L
        d.c = self.L.cassie_sim_duplicate(self.c)
        if not d.c:
            raise RuntimeError('cassie_sim_duplicate failed: ' + _last_error())
        d.nq, d.nv = self.nq, self.nv
        return d

    def copy(self, src):
        self.L.cassie_sim_copy(self.c, src.c)

    # ---- model constants, named as in the reference wrapper (example/cassiemujoco.py:380-610)
    def params(self):
        p = (C.c_int * 6)()
        self.L.cassie_sim_params(self.c, p)
        return list(p)

    def _marr(self, fn, n):
        return np.array(fn(self.c)[:n])

    def get_dof_damping(self, name=None):
        if name:
            return np.array(self.L.cassie_sim_get_dof_name_damping(self.c, name.encode())[:self.L.cassie_sim_get_joint_num_dof(self.c, name.encode())])
        return self._marr(self.L.cassie_sim_dof_damping, self.nv)

    def get_body_mass(self, name=None):
        if name:
            return self.L.cassie_sim_get_body_name_mass(self.c, name.encode())
        return self._marr(self.L.cassie_sim_body_mass, self.params()[4])

    def get_body_ipos(self, name=None):
        if name:
            return np.array(self.L.cassie_sim_get_body_name_ipos(self.c, name.encode())[:3])
        return self._marr(self.L.cassie_sim_body_ipos, 3 * self.params()[4])

    def get_geom_friction(self, name=None):
        if name:
            return np.array(self.L.cassie_sim_get_geom_name_friction(self.c, name.encode())[:3])
        ng = self.params()[5]
        return self._marr(self.L.cassie_sim_geom_friction, 3 * ng).reshape(ng, 3)

    @staticmethod
    def _carr(data):
        a = np.ascontiguousarray(data, dtype=np.float64).ravel()
        return (C.c_double * a.size)(*a)

    def set_dof_damping(self, data, name=None):
        if name:
            self.L.cassie_sim_set_dof_name_damping(self.c, name.encode(), self._carr(np.atleast_1d(data)))
        else:
            assert len(data) == self.nv
            self.L.cassie_sim_set_dof_damping(self.c, self._carr(data))

    def set_body_mass(self, data, name=None):
        if name is None:
            assert len(data) == self.params()[4]
            self.L.cassie_sim_set_body_mass(self.c, self._carr(data))
        else:
            self.L.cassie_sim_set_body_name_mass(self.c, name.encode(), float(data))

    def set_body_ipos(self, data, name=None):
        if name:
            assert len(data) == 3
            self.L.cassie_sim_set_body_name_ipos(self.c, name.encode(), self._carr(data))
        else:
            assert len(data) == 3 * self.params()[4]
            self.L.cassie_sim_set_body_ipos(self.c, self._carr(data))

    def set_geom_friction(self, data, name=None):
        if name is None:
            assert np.size(data) == 3 * self.params()[5]
            self.L.cassie_sim_set_geom_friction(self.c, self._carr(data))
        else:
            assert len(data) == 3
            self.L.cassie_sim_set_geom_name_friction(self.c, name.encode(), self._carr(data))

    def set_const(self):
        self.L.cassie_sim_set_const(self.c)

    def just_set_const(self):
        self.L.cassie_sim_just_set_const(self.c)

    # ---- read-only queries, named as in the reference wrapper (example/cassiemujoco.py:214-306, 815-819)
    def _vec(self, fn, n):
        a = (C.c_double * n)()
        fn(self.c, a)
        return np.array(a[:])

    def foot_forces_raw(self):
        return self._vec(self.L.cassie_sim_foot_forces, 12)

    def get_foot_forces(self):
        f = self.foot_forces_raw()
        return float(np.sqrt((f[0:3] ** 2).sum())), float(np.sqrt((f[6:9] ** 2).sum()))

    def get_heeltoe_forces(self):
        t, h = (C.c_double * 6)(), (C.c_double * 6)()
        self.L.cassie_sim_heeltoe_forces(self.c, t, h)
        return np.array(t[:]), np.array(h[:])

    def check_collision(self, geom_group):
        return bool(self.L.cassie_sim_geom_collision(self.c, int(geom_group)))

    def foot_pos(self):
        return list(self._vec(self.L.cassie_sim_foot_positions, 6))

    def foot_vel(self, vel):
        vel[:12] = self._vec(self.L.cassie_sim_foot_velocities, 12)

    def center_of_mass_position(self):
        return list(self._vec(self.L.cassie_sim_cm_position, 3))

    def center_of_mass_velocity(self):
        return list(self._vec(self.L.cassie_sim_cm_velocity, 3))

    def angular_momentum(self):
        return list(self._vec(self.L.cassie_sim_angular_momentum, 3))

    def check_self_collision(self):
        return bool(self.L.cassie_sim_check_self_collision(self.c))

    def check_obstacle_collision(self):
        return bool(self.L.cassie_sim_check_obstacle_collision(self.c))

    def __del__(self):
        try:
            if getattr(self, 'c', None):
                self.L.cassie_sim_free(self.c)
                self.c = None
        except Exception:
            pass


# ---------------------------------------------------------------- multi-GPU: environments are independent, shard them by index
def env_shard(n_total, rank, world):
    """contiguous env-index range [start, start+count) owned by `rank`: sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def gather_observations(local_obs, group=None):
    """the one optional collective of the path: all-gather of every rank's [n_local, OBS_WIDTH] observation block
    (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).  Ranks may own different env counts."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=local_obs.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local_obs.shape[0]], dtype=torch.int64, device=local_obs.device), group=group)
    counts = [int(c.item()) for c in counts]
    if len(set(counts)) == 1:
        out = torch.empty((world * counts[0],) + tuple(local_obs.shape[1:]), dtype=local_obs.dtype, device=local_obs.device)
        dist.all_gather_into_tensor(out, local_obs.contiguous(), group=group)
        return out
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local_obs.shape[1:]), dtype=local_obs.dtype, device=local_obs.device)
    pad[:local_obs.shape[0]] = local_obs
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
