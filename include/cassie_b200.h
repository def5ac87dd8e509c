/* cassie_b200.h -- C-ABI of the B200-native batched Cassie stepper (libcassie_b200.so).
 *
 * Drop-in boundary for ONE path of osudrl/cassie-mujoco-sim: cassie_sim_step_pd and the few lifecycle / state verbs a caller
 * needs around it.  Every entry point cites the reference interface it replaces (paths relative to /root/reference).
 * Plain pointers and sizes only; no CUDA or torch types appear in any signature (device pointers travel as void*).
 *
 * Two groups:
 *   (1) legacy single-environment verbs with the reference's exact names and signatures (include/cassiemujoco.h) --
 *       a cassie_sim_t is a batch of one environment stepped by the same CUDA kernels (there is no CPU backend);
 *   (2) the new batched verbs (cassie_batch_*, cassie_sim_step_pd_batch) named by BASELINE.json's north_star.
 */
#ifndef CASSIE_B200_H
#define CASSIE_B200_H
#include <stdbool.h>
#include "cassie_bus.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cassie_sim cassie_sim_t;     /* include/cassiemujoco.h:31 */
typedef struct cassie_state cassie_state_t; /* include/cassiemujoco.h:32 */
typedef struct cassie_batch cassie_batch_t; /* new */

/* ------------------------------------------------------------------ (1) legacy single-environment verbs */
/* include/cassiemujoco.h:45  (src/cassiemujoco.c:820-879): loads and caches the model; here: compiles the MJCF (or .cmodel) */
bool cassie_mujoco_init(const char *modelfile);
/* include/cassiemujoco.h:49  (src/cassiemujoco.c:881-...): drops the cached model */
void cassie_cleanup(void);
/* include/cassiemujoco.h:67  (src/cassiemujoco.c:979-1036): NULL on failure, message on stderr */
cassie_sim_t *cassie_sim_init(const char *modelfile, bool reinit);
/* include/cassiemujoco.h:77  (src/cassiemujoco.c:1102-1113): NULL is a no-op */
void cassie_sim_free(cassie_sim_t *sim);
/* include/cassiemujoco.h:95  (src/cassiemujoco.c:1147-1157): THE hot path, one environment, one 0.5 ms tick */
void cassie_sim_step_pd(cassie_sim_t *sim, state_out_t *y, const pd_in_t *u);
/* include/cassiemujoco.h:104,147,183 (src/cassiemujoco.c:1191-1214): borrowed read-write pointers (35 / 32 / 1 doubles) into a
 * host mirror; writes are uploaded before the next step, the mirror is refreshed after every step */
double *cassie_sim_time(cassie_sim_t *sim);
double *cassie_sim_qpos(cassie_sim_t *sim);
double *cassie_sim_qvel(cassie_sim_t *sim);
/* src/cassiemujoco.c:1038-1052 */
int cassie_sim_nv(const cassie_sim_t *sim);
int cassie_sim_nq(const cassie_sim_t *sim);
/* include/cassiemujoco.h:263 (src/cassiemujoco.c:1963-1967): xfrc = force xyz, torque xyz in the world frame; unknown name: no-op */
/* LIMIT of this library: an environment carries ONE (wrench, body) slot.  A call that names another body REPLACES the earlier wrench (the reference
 * keeps xfrc_applied per body, so several bodies can be pushed at once); cassie_sim_clear_forces empties the slot.  The BASELINE configurations push the
 * pelvis only. */
void cassie_sim_apply_force(cassie_sim_t *sim, double xfrc[6], const char *name);
/* include/cassiemujoco.h:268 (src/cassiemujoco.c:1969-1972) */
void cassie_sim_clear_forces(cassie_sim_t *sim);
/* include/cassiemujoco.h:281 (src/cassiemujoco.c:2008-2033) */
void cassie_sim_full_reset(cassie_sim_t *sim);
/* sizes the reference's Python wrapper asks for (src/cassiemujoco.c:1038-1060) and qpos += h qvel (src/cassiemujoco.c:1183-1189; *y is zeroed: the
 * reference runs its estimator on an uninitialised cassie_out_t there) */
int cassie_sim_nbody(const cassie_sim_t *sim);
int cassie_sim_ngeom(const cassie_sim_t *sim);
int cassie_sim_njnt(const cassie_sim_t *sim);
int cassie_sim_nu(const cassie_sim_t *sim);
void cassie_integrate_pos(cassie_sim_t *sim, state_out_t *y);
/* ---- the verbs RL wrappers call around the hot path
 * include/cassiemujoco.h:85 (src/cassiemujoco.c:1137-1145): user torques -> safety layer -> motor model -> physics; *y = the bus of this tick.
 * cassie_sim_step_ethercat (below the safety layer) stays a stub: the kernel always runs cassie_core_sim_step. */
void cassie_sim_step(cassie_sim_t *sim, cassie_out_t *y, const cassie_user_in_t *u);
/* src/cassiemujoco.c:1159-1181: cassie_sim_step_pd with ONE physics step per call whatever the model's timestep */
void cassie_sim_step_pd_no2khz(cassie_sim_t *sim, state_out_t *y, const pd_in_t *u);
/* src/cassiemujoco.c:2090-2092: the bus as the last step's sensors filled it (dynamic subset; the rest as cassie_out_init leaves it, :672-734) */
cassie_out_t cassie_sim_get_cassie_out(cassie_sim_t *sim);
/* include/cassiemujoco.h:108-112 (src/cassiemujoco.c:1196-1204): the physics timestep; control ticks stay 0.5 ms = round(5e-4 / timestep) sub-steps.
 * The pointer is a host mirror: a write takes effect at the next step / forward.  cassie_batch_set_timestep: the same for a batch (0 / -1). */
double *cassie_sim_timestep(cassie_sim_t *sim);
void cassie_sim_set_timestep(cassie_sim_t *sim, double dt);
int cassie_batch_set_timestep(cassie_batch_t *b, double dt);
/* src/cassiemujoco.c:1221-1225: mj_forward on the current state; returns 0 */
int cassie_sim_forward(cassie_sim_t *sim);
/* include/cassiemujoco.h:271-275 (src/cassiemujoco.c:1974-2000): pin / free the pelvis (stiff spring-damper on its slides, damping on its ball joint) */
void cassie_sim_hold(cassie_sim_t *sim);
void cassie_sim_release(cassie_sim_t *sim);
/* src/cassiemujoco.c:1466-1541 (example/test_terrain.c:118-157 builds stairs out of the 15 boxes of model/cassie.xml:232-246 with them): geom position /
 * orientation / size in the reference's geom numbering; borrowed pointers are host mirrors, a change takes effect at the next step / forward.
 * All 135 box x robot candidate pairs are collided (contact rules for boxes: DESIGN.md section 3); a box out of the robot's reach costs one
 * distance test per step.  cassie_batch_set_geom_pose: the same for a batch, one placement shared by its environments (NULL: leave that part). */
double *cassie_sim_geom_pos(cassie_sim_t *sim);
double *cassie_sim_geom_quat(cassie_sim_t *sim);
double *cassie_sim_geom_size(cassie_sim_t *sim);
double *cassie_sim_geom_name_pos(cassie_sim_t *sim, const char *name);
double *cassie_sim_geom_name_quat(cassie_sim_t *sim, const char *name);
double *cassie_sim_geom_name_size(cassie_sim_t *sim, const char *name);
void cassie_sim_set_geom_pos(cassie_sim_t *sim, double *pos);
void cassie_sim_set_geom_quat(cassie_sim_t *sim, double *quat);
void cassie_sim_set_geom_size(cassie_sim_t *sim, double *size);
void cassie_sim_set_geom_name_pos(cassie_sim_t *sim, const char *name, double *pos);
void cassie_sim_set_geom_name_quat(cassie_sim_t *sim, const char *name, double *quat);
void cassie_sim_set_geom_name_size(cassie_sim_t *sim, const char *name, double *size);
int cassie_batch_set_geom_pose(cassie_batch_t *b, const char *name, const double *pos, const double *quat, const double *size);
/* include/cassiemujoco.h:61-73 (src/cassiemujoco.c:1072-1093): model (per-environment constants, height field, timestep) + full dynamic state */
cassie_sim_t *cassie_sim_duplicate(const cassie_sim_t *src);
void cassie_sim_copy(cassie_sim_t *dst, const cassie_sim_t *src);
/* include/cassiemujoco.h:434-463 (src/cassiemujoco.c:3380-3452): full dynamic state of a simulator -- qpos, qvel, warm start, time, sensor snapshot,
 * cassie_out subset, encoder filters, torque delay line, applied forces, estimator filters, last observation row.  The borrowed time / qpos / qvel
 * pointers are host mirrors: what the caller writes there after cassie_get_state is honoured by cassie_set_state.  A state owns its (device)
 * memory: it can be restored into any simulator of the same model and freed at any time, before or after the simulator it was taken from. */
cassie_state_t *cassie_state_alloc(void);
cassie_state_t *cassie_state_duplicate(const cassie_state_t *src);
void cassie_state_copy(cassie_state_t *dst, const cassie_state_t *src);
void cassie_state_free(cassie_state_t *state);
double *cassie_state_time(cassie_state_t *state);
double *cassie_state_qpos(cassie_state_t *state);
double *cassie_state_qvel(cassie_state_t *state);
void cassie_get_state(const cassie_sim_t *sim, cassie_state_t *state);
void cassie_set_state(cassie_sim_t *sim, const cassie_state_t *state);
/* Import compatibility: the library also exports, as stubs, the 133 further names example/cassiemujoco_ctypes.py resolves at import time
 * (cassie_vis_*, UDP, pack / unpack, pd_input_* / cassie_core_sim_* / state_output_* host objects, mjModel / mjData accessors, ...;
 * csrc/legacy_stubs.inc, list in tests/golden/ctypes_bound_names.txt).  They are outside the accelerated path: a call records an error
 * (cassie_b200_last_error), prints it and returns 0 / NULL.  They are deliberately not declared here. */
/* include/cassiemujoco.h:317-329 (src/cassiemujoco.c:2050-2080): height-field terrain of cassie_hfield.xml; nrow*ncol floats in [0,1],
 * row-major, row <-> y, column <-> x.  cassie_sim_hfielddata returns a borrowed read-write host mirror uploaded before the next step. */
int cassie_sim_get_hfield_nrow(cassie_sim_t *sim);
int cassie_sim_get_hfield_ncol(cassie_sim_t *sim);
int cassie_sim_get_nhfielddata(cassie_sim_t *sim);
float *cassie_sim_hfielddata(cassie_sim_t *sim);
void cassie_sim_set_hfielddata(cassie_sim_t *sim, float *data);
/* read-only derived quantities.  include/cassiemujoco.h:200-240 (src/cassiemujoco.c:1586-1699, 1812-1898, 1944-1961).
 * Like the reference, the contact / foot group reads what the LAST step left behind (contact list and forces of the state that step
 * started from); the centre-of-mass group first recomputes the kinematics of the current state (the reference calls mj_fwdPosition).
 * cm_velocity / angular_momentum use velocities consistent with that state (the reference mixes them with the previous step's cvel). */
bool cassie_sim_check_obstacle_collision(const cassie_sim_t *sim);
bool cassie_sim_check_self_collision(const cassie_sim_t *sim);
void cassie_sim_foot_forces(const cassie_sim_t *c, double cfrc[12]);
void cassie_sim_heeltoe_forces(const cassie_sim_t *c, double toe_force[6], double heel_force[6]);
bool cassie_sim_geom_collision(const cassie_sim_t *c, int geom_group);
void cassie_sim_foot_positions(const cassie_sim_t *c, double cpos[6]);   /* src/cassiemujoco.c:1608-1621 (defined there, not in the header) */
void cassie_sim_foot_velocities(const cassie_sim_t *c, double cvel[12]);
void cassie_sim_cm_position(const cassie_sim_t *c, double cm_pos[3]);
void cassie_sim_cm_velocity(const cassie_sim_t *c, double cm_vel[3]);
void cassie_sim_angular_momentum(const cassie_sim_t *c, double Lcm[3]);
/* model constants for domain randomisation.  src/cassiemujoco.c:1303-1436, 949-977 (most are defined there without a header
 * declaration; example/cassiemujoco.py:517-610 calls them).  The pointer-returning verbs hand out borrowed read-write host mirrors
 * ([nv], [nbody], [3 nbody], [3 ngeom] doubles, numbered like the reference's mjModel incl. its visual mesh geoms); writes are uploaded
 * before the next step / query / set_const.  What the stepper uses: dof_damping of the 32 robot dofs, body_mass, body_ipos (not of the
 * free `cup_box` body), the sliding coefficient geom_friction[3 g] of colliding geoms.  As in MuJoCo, a changed mass acts at once while
 * the solver's reference weights (body/dof_invweight0, meaninertia, subtree masses) stay stale until (just_)set_const. */
void cassie_sim_params(cassie_sim_t *c, int *params);                 /* :1566-1574: nq nv nu nsensordata nbody ngeom */
double *cassie_sim_dof_damping(cassie_sim_t *c);                       /* :1303 */
double *cassie_sim_body_mass(cassie_sim_t *c);                         /* :1308 */
double *cassie_sim_body_ipos(cassie_sim_t *c);                         /* :1313 */
double *cassie_sim_geom_friction(cassie_sim_t *c);                     /* :1318 */
void cassie_sim_set_dof_damping(cassie_sim_t *c, double *damp);        /* :1330 */
void cassie_sim_set_dof_name_damping(cassie_sim_t *c, const char *name, double *damp);   /* :1338 */
double *cassie_sim_get_dof_name_damping(cassie_sim_t *c, const char *name);              /* :1347 */
int cassie_sim_get_joint_num_dof(cassie_sim_t *c, const char *name);                     /* :1353 */
void cassie_sim_set_body_mass(cassie_sim_t *c, double *mass);          /* :1366 */
void cassie_sim_set_body_name_mass(cassie_sim_t *c, const char *name, double mass);      /* :1373 */
double cassie_sim_get_body_name_mass(cassie_sim_t *c, const char *name);                 /* :1379 */
void cassie_sim_set_body_ipos(cassie_sim_t *c, double *ipos);          /* :1385 (reads ipos[i + j], as the reference does) */
void cassie_sim_set_body_name_ipos(cassie_sim_t *c, const char *name, double *ipos);     /* :1394 */
double *cassie_sim_get_body_name_ipos(cassie_sim_t *c, const char *name);                /* :1402 */
void cassie_sim_set_geom_friction(cassie_sim_t *c, double *fric);      /* :1420 */
void cassie_sim_set_geom_name_friction(cassie_sim_t *c, const char *name, double *fric); /* :1427 (addresses the geom's own triple) */
double *cassie_sim_get_geom_name_friction(cassie_sim_t *c, const char *name);            /* :1433 */
void cassie_sim_set_const(cassie_sim_t *c);                            /* :949-972: mj_setConst, then qpos <- init, qvel <- 0, time <- 0, mj_forward */
void cassie_sim_just_set_const(cassie_sim_t *c);                       /* :974-977: mj_setConst only */
/* src/cassiemujoco.c:2002-2006: the 16 radio channels; channel 8 < 1 engages safe-torque-off */
void cassie_sim_radio(cassie_sim_t *sim, double channels[16]);

/* ------------------------------------------------------------------ (2) batched verbs (new; north_star) */
#define CASSIE_B200_FP32 0 /* throughput build: state and arithmetic in fp32 */
#define CASSIE_B200_FP64 1 /* parity build: state and arithmetic in fp64 */
#define CASSIE_PD_WIDTH 52  /* compact motor-PD row: torque[10] pTarget[10] dTarget[10] pGain[10] dGain[10] pad[2] */
#define CASSIE_OBS_WIDTH 112 /* compact observation row, see cassie_batch_get_obs */
#define CASSIE_AUX_WIDTH 64 /* derived-quantity row, see cassie_batch_get_aux */
/* offsets inside a derived-quantity row */
#define CASSIE_AUX_FOOT_FORCE 0   /* [12] cassie_sim_foot_forces layout: left xyz, 3 zeros, right xyz, 3 zeros */
#define CASSIE_AUX_TOE_FORCE 12   /* [6]  cassie_sim_heeltoe_forces toe_force: left xyz, right xyz */
#define CASSIE_AUX_HEEL_FORCE 18  /* [6]  ... heel_force */
#define CASSIE_AUX_FOOT_POS 24    /* [6]  cassie_sim_foot_positions */
#define CASSIE_AUX_FOOT_VEL 30    /* [12] cassie_sim_foot_velocities */
#define CASSIE_AUX_CM_POS 42      /* [3]  centre of mass of everything (cassie_sim_cm_position) */
#define CASSIE_AUX_CM_VEL 45      /* [3]  its velocity (cassie_sim_cm_velocity) */
#define CASSIE_AUX_ANGMOM 48      /* [3]  angular momentum about it (cassie_sim_angular_momentum) */
#define CASSIE_AUX_OBSTACLE 51    /* 1 if a contact involves an obstacle geom (cassie_sim_check_obstacle_collision) */
#define CASSIE_AUX_SELF 52        /* 1 if two robot geoms touch (cassie_sim_check_self_collision) */
#define CASSIE_AUX_GROUPMASK 53   /* bit g set: a group-1 geom touches a geom of group g (cassie_sim_geom_collision) */
#define CASSIE_AUX_NCON 54        /* contacts that entered the solve */

/* n_env environments on CUDA device `device`, all in the state cassie_sim_init leaves (src/cassiemujoco.c:979-1036).
 * modelfile: MJCF (.xml) or a compiled table (.cmodel).  NULL + stderr message on failure (no GPU, bad model, ...). */
cassie_batch_t *cassie_batch_init(const char *modelfile, int n_env, int device, int precision);
void cassie_batch_free(cassie_batch_t *b);
int cassie_batch_nenv(const cassie_batch_t *b);
int cassie_batch_nq(const cassie_batch_t *b);
int cassie_batch_nv(const cassie_batch_t *b);
/* back to the cassie_sim_init state (mask == NULL: all envs; else mask[i] != 0 selects env i) */
void cassie_batch_reset(cassie_batch_t *b, const unsigned char *mask);

/* cassie_sim_step_pd for every environment: pd_in[n_env] host AoS in, state_out[n_env] host AoS out (may be NULL).
 * Replaces n_env calls of src/cassiemujoco.c:1147-1157.  Synchronous.  state_out carries the stateless part of the reference's
 * state_output_step, decoded from the closed archive and checked equal to it to 1e-11: motor / joint position + velocity + torque,
 * pelvis.orientation (+-q, mat2quat's sign), rotationalVelocity, translationalAcceleration, both feet's position / orientation (pelvis frame) and
 * footRotationalVelocity / footTranslationalVelocity (foot frame), radio, battery; toeForce / heelForce when enabled (see
 * cassie_batch_enable_estimator_forces); pelvis.position / translationalVelocity / externalForce and terrain.height (the estimator's filters)
 * when enabled (see cassie_batch_enable_estimator_filter).  externalMoment and terrain.slope are zero, as they are in the reference's output. */
void cassie_sim_step_pd_batch(cassie_batch_t *envs, const pd_in_t *pd_in, state_out_t *state_out);

/* throughput path: compact rows.  pd: host [n_env][CASSIE_PD_WIDTH] doubles, copied to the device (and converted to the batch
 * precision) on the batch stream; step: `nticks` control ticks per launch with the PD rows held; asynchronous. */
void cassie_batch_set_pd(cassie_batch_t *b, const double *pd);
/* the taskPd branch of pd_in_t (include/pd_in_t.h:32-38) for the throughput path: host [n_env][60] doubles = per leg (left, right)
 * torque[6] pTarget[6] dTarget[6] pGain[6] dGain[6]; NULL switches the branch off.  Task coordinates as the reference's closed
 * pd_input_step uses them (decoded, pinned to the archive): foot position in the pelvis frame, then yaw / pitch / roll of the foot
 * frame; rates in the foot frame.  cassie_sim_step_pd(_batch) forward pd_in_t's taskPd fields automatically.  0 / -1. */
int cassie_batch_set_task_pd(cassie_batch_t *b, const double *rows);
/* Open-loop gait on top of the motor-PD rows (BASELINE config 5 "random PD gaits"): in every control tick of every launch
 * pTarget_i(t) = row.pTarget_i + amp_i sin(2 pi freq t + phase_i), t = control ticks since the last reset (or since this call) x 0.5 ms.
 * amp [n][10], freq [n] (Hz), phase [n][10]; all NULL switches it off.  New verb: the reference moves its targets from the host every tick. */
int cassie_batch_set_pd_gait(cassie_batch_t *b, const double *amp, const double *freq, const double *phase);
void cassie_batch_step(cassie_batch_t *b, int nticks);
void cassie_batch_sync(cassie_batch_t *b);
/* host copies (synchronous): qpos [n][35], qvel [n][32], time [n], obs [n][CASSIE_OBS_WIDTH] =
 * motor pos[10] vel[10] torque[10], joint pos[6] vel[6], IMU quat[4] gyro[3] accel[3] mag[3], time (0..55);
 * estimator: translationalAcceleration[3] (56), pad, per foot {position 3, orientation 4, rotational velocity 3, translational
 * velocity 3} left (60..72) right (73..85), pelvis.orientation[4] (86..89; the IMU quaternion up to sign), pad;
 * in-kernel estimator (cassie_batch_enable_estimator_device; zero while it is off): pelvis.position[3] (96), translationalVelocity[3] (99),
 * externalForce[3] (102), terrain.height (105), toeForce = heelForce left[3] (106) right[3] (109) */
void cassie_batch_get_qpos(cassie_batch_t *b, double *out);
void cassie_batch_set_qpos(cassie_batch_t *b, const double *in);
void cassie_batch_get_qvel(cassie_batch_t *b, double *out);
void cassie_batch_set_qvel(cassie_batch_t *b, const double *in);
void cassie_batch_get_time(cassie_batch_t *b, double *out);
void cassie_batch_get_obs(cassie_batch_t *b, double *out);
/* derived quantities for every environment, as by-products of the step kernel (no second pass over the state).
 * enable_aux(b, 1) allocates [n][CASSIE_AUX_WIDTH] rows in the batch precision; from then on every step / forward launch fills them:
 * contact-derived slots and foot positions / velocities exactly as the reference's queries would return right after that step; the
 * centre-of-mass slots describe the state the last sub-step STARTED from (one 0.5 ms tick old).  cassie_batch_query recomputes only the
 * centre-of-mass slots for the CURRENT state and writes nothing else (state, sensors and the other slots are untouched).
 * get_aux: synchronous host copy [n][CASSIE_AUX_WIDTH] doubles; device pointer: cassie_batch_device_ptr(b, "aux").  0 / -1. */
int cassie_batch_enable_aux(cassie_batch_t *b, int on);
int cassie_batch_get_aux(cassie_batch_t *b, double *out);
int cassie_batch_query(cassie_batch_t *b);
/* per-environment model constants (domain randomisation).  rows are host [n][width] doubles in the host model's numbering (= the
 * reference's mjModel): body_mass [nbody], body_ipos [3 nbody], dof_damping [nv], geom_friction [3 ngeom].  The first call allocates a
 * 1 KB (fp32) constant row per environment that the step kernel then reads instead of the shared model block.  set_const runs mj_setConst
 * for the selected environments (mask == NULL: all) as one kernel launch at the reference configuration -- body/dof inverse weights, subtree
 * masses, mean inertia -- and, with reset_state != 0, the state reset of cassie_sim_set_const (src/cassiemujoco.c:955-971).  0 / -1. */
int cassie_batch_nbody(const cassie_batch_t *b);
int cassie_batch_ngeom(const cassie_batch_t *b);
int cassie_batch_set_body_mass(cassie_batch_t *b, const double *mass);
int cassie_batch_set_body_ipos(cassie_batch_t *b, const double *ipos);
int cassie_batch_set_dof_damping(cassie_batch_t *b, const double *damp);
int cassie_batch_set_geom_friction(cassie_batch_t *b, const double *fric);
int cassie_batch_get_body_mass(cassie_batch_t *b, double *mass);
int cassie_batch_get_body_ipos(cassie_batch_t *b, double *ipos);
int cassie_batch_get_dof_damping(cassie_batch_t *b, double *damp);
int cassie_batch_get_geom_friction(cassie_batch_t *b, double *fric);
int cassie_batch_set_const(cassie_batch_t *b, const unsigned char *mask, int reset_state);
/* toeForce / heelForce of state_out_t (the estimator's spring-force model, decoded; single-precision agreement with the archive): computed
 * on the host while cassie_sim_step_pd_batch unpacks its rows; off by default for batches (host time per environment), always on for a
 * cassie_sim_t.  The pure function behind it is exported for callers that keep observations on the device:
 * ang = hipRoll, hipYaw, hipPitch, knee (motor positions), shin, tarsus (joint encoders), foot (motor position); quat = IMU quaternion. */
int cassie_batch_enable_estimator_forces(cassie_batch_t *b, int on);
void cassie_b200_estimator_leg_force(int side, const double ang[7], const double quat[4], double force[3]);
/* The estimator's filters (state_output_step's stateful part, src/cassiemujoco.c:1180; decoded from the closed block's memory, agreement with the
 * archive 1e-13 on the archive's own stateless outputs): pelvis.position, translationalVelocity, externalForce, terrain.height.  One filter
 * object per environment lives on the host and advances once per cassie_sim_step_pd_batch call (the reference runs it once per 2 kHz
 * step_pd call); it turns the force option on.  Off by default for batches, always on for a cassie_sim_t.  cassie_batch_reset_estimator
 * restarts the filters (state_output_setup; mask as in cassie_batch_reset).  The filter is also exported as a plain host object:
 * ..._step reads orientation, translationalAcceleration, both feet's position and toe / heel forces from *y and writes the four outputs. */
int cassie_batch_enable_estimator_filter(cassie_batch_t *b, int on);
/* The same estimator inside the step kernel (leg forces + filters, every 2 kHz tick of a launch, so multi-tick launches keep it exact and
 * device-resident observations are complete): columns 96..111 of the observation row ("obs"), also copied out by cassie_batch_get_estimator --
 * pelvis.position 3, translationalVelocity 3, externalForce 3, terrain.height 1, toeForce (= heelForce) left 3, right 3.
 * The filter state is held in doubles in every batch precision.  cassie_batch_reset(_estimator) restarts it.
 * cassie_sim_step_pd_batch switches it on by itself at the first call that asks for state_out_t rows, because cassie_sim_step_pd runs
 * state_output_step in every call (src/cassiemujoco.c:1156); cassie_batch_enable_estimator_device(b, 0) keeps it off (the filtered fields,
 * toeForce and heelForce of state_out_t are then zero).  A cassie_sim_t always runs it. */
#define CASSIE_EST_WIDTH 16
int cassie_batch_enable_estimator_device(cassie_batch_t *b, int on);
int cassie_batch_get_estimator(cassie_batch_t *b, double *out /* [n][CASSIE_EST_WIDTH] */);
int cassie_batch_reset_estimator(cassie_batch_t *b, const unsigned char *mask);
void *cassie_b200_estimator_filter_new(void);
void cassie_b200_estimator_filter_free(void *filter);
void cassie_b200_estimator_filter_reset(void *filter);
void cassie_b200_estimator_filter_step(void *filter, state_out_t *y);
/* measurement aids: accumulated wall-clock seconds of cassie_sim_step_pd_batch since the last reset -- out[0] host pack of pd_in_t[], out[1] device part
 * (H2D, kernel, D2H, as waited for), out[2] host unpack into state_out_t[], out[3] number of calls, out[4] / out[5] milliseconds of the H2D copy /
 * of the kernel by CUDA events (only when the environment variable CASSIE_B200_AOS_EVENTS is set); CPUs this process may use (affinity mask
 * clipped by the cgroup quota), which bounds the pack / unpack threads (CASSIE_B200_AOS_THREADS, default 32) */
void cassie_batch_aos_timing(cassie_batch_t *b, double out[6], int reset);
int cassie_b200_effective_cpus(void);
/* OpenMP threads cassie_sim_step_pd_batch packs / unpacks with: CASSIE_B200_AOS_THREADS, else the CPUs above minus two (at most 32) */
int cassie_b200_aos_threads(void);
/* Batched snapshots (new): every row array of the batch copied device-to-device into an opaque handle; restore all environments or the masked
 * subset (mask as in cassie_batch_reset; e.g. return the fallen robots of an RL batch to a stored standing state).  0 / -1. */
void *cassie_batch_state_alloc(cassie_batch_t *b);
void cassie_batch_state_free(cassie_batch_t *b, void *state);
int cassie_batch_get_state(cassie_batch_t *b, void *state);
int cassie_batch_set_state(cassie_batch_t *b, const void *state, const unsigned char *mask);
/* re-run mj_forward on the current state (after set_qpos / set_qvel), like cassie_sim_forward (src/cassiemujoco.c:1221-1225) */
void cassie_batch_forward(cassie_batch_t *b);
/* batched cassie_sim_apply_force / cassie_sim_clear_forces: xfrc [n][6]; one perturbed body per env */
int cassie_batch_apply_force(cassie_batch_t *b, const double *xfrc, const char *body_name);
void cassie_batch_clear_forces(cassie_batch_t *b);
/* batched cassie_integrate_pos (src/cassiemujoco.c:1183-1189 -> mj_integratePos): qpos <- qpos (+) h * qvel, the HBM-bound kernel */
void cassie_batch_integrate_pos(cassie_batch_t *b);

/* batched cassie_sim_set_hfielddata: `n_terrains` height fields of nrow*ncol floats; environment e stands on terrain e % n_terrains
 * (BASELINE config 4: a few dozen terrains shared round-robin).  0 on success, -1 if the model has no height field. */
int cassie_batch_set_hfielddata(cassie_batch_t *b, const float *data, int n_terrains);
int cassie_batch_hfield_nrow(const cassie_batch_t *b);
int cassie_batch_hfield_ncol(const cassie_batch_t *b);

/* zero-copy access for a PyTorch / DLPack caller: device pointer of a state array ("qpos" [n][36], "qvel" [n][32],
 * "pd" [n][52], "obs" [n][112], "xfrc" [n][8], "aux" [n][64]) in the batch precision; the stream all work is enqueued on (cudaStream_t). */
void *cassie_batch_device_ptr(cassie_batch_t *b, const char *field);
void cassie_batch_set_stream(cassie_batch_t *b, void *cuda_stream);
void *cassie_batch_get_stream(cassie_batch_t *b);
int cassie_batch_precision(const cassie_batch_t *b);
/* row width (in elements) of a device array named as in cassie_batch_device_ptr: qpos 36 (44 with the extra free body of cassie_tray_box.xml), qvel 32 (40) */
int cassie_batch_row_width(const cassie_batch_t *b, const char *field);
/* per-env solver statistics of the last sub-step: int [n][8] = nefc, ncon, nlimit, PGS iterations, dropped contacts, 0,0,0 */
void cassie_batch_get_counters(cassie_batch_t *b, int *out);
/* launches issued since init (bench.py's gpu_launches) */
long cassie_batch_launch_count(const cassie_batch_t *b);
/* test hook: stage-by-stage intermediates of env `env` from the last sub-step (DevModel debug layout, doubles) */
int cassie_batch_debug_dump(cassie_batch_t *b, int env, double *out, int n);
/* last error message of this thread ("" if none) */
const char *cassie_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CASSIE_B200_H */
