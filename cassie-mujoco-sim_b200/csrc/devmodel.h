// devmodel.h -- the per-model constant block the stepper reads (one copy per CTA in shared memory, staged with a
// TMA bulk copy), the HBM state layout, and the per-warp scratch layout.  Shared by the CUDA build (kernels.cu) and
// by the test-only host emulation of the same source (tests/emu/).
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#define CASSIE_HD __host__ __device__
#else
#define CASSIE_HD
#endif

namespace cassie {

constexpr int MB = 32;        // bodies (nbody <= 32: one lane per body)
constexpr int MJ = 32;        // joints
constexpr int MV = 32;        // dofs   (nv <= 32: one lane per dof)
constexpr int MG = 16;        // collision geoms on moving bodies (world pose recomputed every step)
constexpr int MGS = 16;       // static collision geoms (world body or bodies welded to it: floor plane, height field, the 15 stair boxes of cassie.xml): device ids MG .. MG + MGS - 1
constexpr int MGT = MG + MGS;
constexpr int MPAIR = 160;    // candidate geom pairs in MuJoCo's order (cassie.xml: 9 floor + 9 x 15 box + 9 leg-leg = 153), 32 per collision pass
constexpr int NPC = 8;        // distinct contact-parameter records among the pairs (mj_contactParam results)
constexpr int ME = 4;         // connect equalities
constexpr int MU = 10;        // motors
constexpr int NM_MAX = 320;   // sparse mass-matrix entries (307 for Cassie)
constexpr int NTRI_MAX = 276; // off-diagonal entries of the sparse factor (275 for Cassie)
constexpr int NFAC_MAX = 1520; // rank-1 update pairs of the factorisation schedule (1519 for Cassie)
constexpr int NEFC = 48;      // constraint rows per env (12 equality + limits + 4 per floor contact); excess contacts are dropped and counted
constexpr int MAXCON = 12;    // contacts per env
constexpr int YSTRIDE_MAIN = 33;  // row stride of the constraint matrix in shared memory: dofs + 1 (odd: bank-conflict free both ways)
constexpr int YSTRIDE_MAX = 39;   // with the 6 dofs of an extra free body (cassie_tray_box.xml)

// model features a kernel instance is compiled for (template parameter FEAT): an instance without a feature carries none of its code
constexpr int F_XB = 1, F_HFIELD = 2, F_BOX = 4, F_ALL = 7;   // extra free body (cassie_tray_box.xml's cup), height field, box geoms

CASSIE_HD inline int pair_g1(uint32_t c) { return (int)(c & 63u); }
CASSIE_HD inline int pair_g2(uint32_t c) { return (int)((c >> 6) & 63u); }
CASSIE_HD inline int pair_kind(uint32_t c) { return (int)((c >> 12) & 15u); }
CASSIE_HD inline int pair_pc(uint32_t c) { return (int)((c >> 16) & 15u); }
CASSIE_HD inline int pair_rank(uint32_t c) { return (int)((c >> 20) & 255u); }

// pair kinds handled by the narrow phase
enum PairKind { PAIR_PLANE_SPHERE = 0, PAIR_PLANE_CAPSULE = 1, PAIR_CAPSULE_CAPSULE = 2, PAIR_HFIELD_SPHERE = 3, PAIR_HFIELD_CAPSULE = 4,
                PAIR_PLANE_BOX = 5, PAIR_SPHERE_BOX = 6, PAIR_CAPSULE_BOX = 7, PAIR_BOX_BOX = 8 };

template <typename real>
struct DevModel {
  // ---- sizes / options
  int qpos_w, qvel_w, ystride, xb;   // HBM row widths, constraint-matrix row stride, id of the extra free body (or -1)
  int xb_qadr, xb_dadr, xb_jnt, any_jnt_pos;   // any_jnt_pos: some joint anchor is not at its body origin
  int nq, nv, nbody, njnt, ngeom, npair, neq, nu, maxdepth, nM, ntri, nsub, iterations, imu_body, has_damping, force_zpath;
  real timestep, tolerance, pgs_scale, root_mass_inv, euler_eps, padr[3];
  real gravity[3], magnetic[3], imu_pos[3], imu_quat[4], imu_mat[9], gyro_cutoff, accel_cutoff;
  int hf_nrow, hf_ncol, padh[2];
  real hf_size[4];           // height field: x half-size, y half-size, elevation scale, base thickness
  real xb_mass, xb_inertia[3], xb_dsqi[6], padxb[2];  // extra free body: mass, principal inertia (inertial frame = body frame), 1/sqrt of its diagonal mass matrix
  // ---- bodies
  int body_parent[MB], body_depth[MB], body_jntadr[MB], body_jntnum[MB], body_lastdof[MB], body_subtree_end[MB];
  uint32_t body_kid_dofs[MB];   // bodies with a large subtree: (first dof + 1) of up to 4 direct children, 8 bits each, 0 = none; their subtree sums are composed from the children's (0: summed directly)
  int any_big, padb[3];
  uint32_t body_dofmask[MB];
  real body_pos[MB][3], body_quat[MB][4], body_ipos[MB][3], body_imat[MB][9], body_mass[MB], body_inertia[MB][3], body_invw[MB];
  // mirror symmetry of the dof tree (two identical legs below a common base chain): lets the row transform and the A = Y Y' products skip the
  // leg a constraint row does not touch.  body_side: 0 base / world / extra body, 1 first leg, 2 second leg
  int sym_on, sym_first, sym_n, sym_madr;
  int spec19, padsp[3];   // the dof tree is the one csrc/cassie_tree_gen.inc was generated for: single-leg rows take the straight-line transform
  unsigned char body_side[MB];
  // ---- joints
  int jnt_type[MJ], jnt_qposadr[MJ], jnt_dofadr[MJ], jnt_body[MJ], jnt_limited[MJ];
  real jnt_pos[MJ][3], jnt_axis[MJ][3], jnt_stiffness[MJ], jnt_range[MJ][2], jnt_qpos0[MJ], jnt_qspring[MJ], jnt_solref[MJ][2], jnt_solimp[MJ][5];
  // ---- dofs
  int dof_body[MV], dof_jnt[MV], dof_parent[MV], dof_Madr[MV], dof_depth[MV], dof_cvelsrc[MV], dof_Mrow[MV], dof_subtree_end[MV];
  unsigned char dof_anc[MV][16];  // t-th ancestor of a dof (t = 1: parent)
  uint32_t dof_ancmask[MV];
  real dof_armature[MV], dof_damping[MV], dof_invweight0[MV];
  uint32_t tri[NTRI_MAX];     // (i << 24) | (j << 16) | qLD address of L(i,j); i descending, ancestors nearest first
  int nfac, fac_start[MV + 1], padf[2];
  uint32_t fac_pairs[NFAC_MAX];  // (t << 24) | (src << 12) | dst: qLD[dst] -= qLD[src] * f_t, grouped by eliminated dof k
  // ---- collision geoms and pairs
  int geom_body[MGT], geom_type[MGT], ngeom_static, root_body, npair_a, static_box_mask, padn[2];   // npair_a: pairs without a static box come first
  real geom_pos[MG][3], geom_mat[MG][9];   // moving geoms: frame in their body (row-major rotation)
  real geom_wpose[MGS][12];                // static geoms: world pose as the collision stage wants it (position, z axis, x axis, y axis)
  real geom_size[MGT][3], geom_rbound[MGT];   // sizes, bounding-sphere radius (0: plane / height field)
  real robot_reach, padg[3];               // no robot collision geom reaches farther than this from the root body's origin (obstacle broad phase)
  // candidate pairs: (g1 | g2 << 6 | kind << 12 | parameter record << 16 | rank in MuJoCo's pair order << 20); records = distinct mj_contactParam
  // results.  Stored in two runs, each in MuJoCo's order: [0, npair_a) the pairs without a static box, [npair_a, npair) the static-box pairs, which are
  // only visited when a box is within the robot's reach; contacts are put back into MuJoCo's order by rank afterwards
  uint32_t pair_code[MPAIR];
  int pc_condim[NPC], pc_mu_src[NPC];   // mu_src: how the sliding friction follows from the geoms': 0 max of both (equal priority), 1 geom 1, 2 geom 2
  int pc_flags[NPC];    // derived-quantity flags: bit 0 obstacle geom involved (geom user == 1), bit 1 robot-robot (both user == 2),
                        // bits 8.. : geom groups g met by a group-1 geom (cassie_sim_geom_collision, src/cassiemujoco.c:1944-1961)
  real pc_mu[NPC], pc_margin[NPC], pc_gap[NPC], pc_solref[NPC][2], pc_solimp[NPC][5];
  real qpos0[44];       // the reference configuration (mj_setConst works there)
  // ---- feet (src/cassiemujoco.c:861-866): body ids, toe / heel points in the foot frames, total mass including the extra free body
  int foot_body[2], padfb[2];
  real toe_local[2][3], heel_local[2][3], foot_offset, total_mass_inv, padft[2];
  // ---- equality
  int eq_b1[ME], eq_b2[ME];
  real eq_data[ME][6], eq_solref[ME][2], eq_solimp[ME][5];
  // ---- motors and encoders (model/cassie.xml:258-287)
  int act_dof[MU], act_qposadr[MU], enc_bits[16], enc_qposadr[16], pad2[2];
  real act_gear[MU], act_ctrl_lo[MU], act_ctrl_hi[MU], act_wmax[MU], act_torque_limit[MU], enc_scale[16];
};

// ---- HBM state: one row per environment in each array (row-major, env index slowest)
constexpr int QPOS_W_MAIN = 36, QVEL_W_MAIN = 32;   // row widths for the nq 35 / nv 32 models (runtime values live in DevModel::qpos_w / qvel_w)
constexpr int QPOS_W_XB = 44, QVEL_W_XB = 40;       // with an extra free body (nq 42 / nv 38)
constexpr int CST_W = 192;      // controller / sensor state, layout below
constexpr int DFILT_W = 96;     // int32 drive FIR taps [10][9], then DF_TICK
constexpr int DF_TICK = 90;     // control ticks since the last reset (int32): the clock of the optional open-loop gait generator
constexpr int GAIT_W = 24;      // optional per-env sinusoidal gait on the motor-PD targets: amplitude[10], phase[10], frequency (Hz), pad
constexpr int GA_AMP = 0, GA_PHASE = 10, GA_FREQ = 20;
constexpr int PD_W = 52;        // torque, pTarget, dTarget, pGain, dGain for the 10 motors (+2 pad)
constexpr int TASK_W = 64;      // optional task-space PD rows: per leg torque, pTarget, dTarget, pGain, dGain [6] each (left 0..29, right 30..59)
constexpr int XFRC_W = 8;       // force xyz, torque xyz, body id (as real), pad
constexpr int OBS_W = 112;      // the dynamic subset of cassie_out_t copied out at src/cassiemujoco.c:1127 + the decoded estimator outputs (stateless part, then the filtered part)
// CST offsets
constexpr int CS_SENSOR = 0;    // sensordata[29]
constexpr int CS_ACTVEL = 32;   // actuator_velocity[10]
constexpr int CS_DPOS = 42, CS_DVEL = 52, CS_DTORQUE = 62;   // cassie_out drive position / velocity / torque
constexpr int CS_JPOS = 72, CS_JVEL = 78;                    // cassie_out joint position / velocity
constexpr int CS_DELAY = 84;    // torque_delay[10][6]
constexpr int CS_JFX = 144;     // joint filter x[6][4]
constexpr int CS_JFY = 168;     // joint filter y[6][3]
constexpr int CS_TIME = 186;
constexpr int CS_STO = 187;     // radio channel 8 (safe-torque-off when < 1)
// OBS offsets
constexpr int OB_MPOS = 0, OB_MVEL = 10, OB_MTORQUE = 20, OB_JPOS = 30, OB_JVEL = 36, OB_QUAT = 42, OB_GYRO = 46, OB_ACCEL = 49, OB_MAG = 52, OB_TIME = 55;
// stateless part of the reference's estimator (state_output_step, closed source; semantics recovered by probing the archive):
constexpr int OB_EST_ACC = 56;    // [3] pelvis.translationalAcceleration
constexpr int OB_FOOT = 60;       // [2][13] per foot: position 3, orientation 4 (pelvis frame), rotational velocity 3, translational velocity 3 (foot frame)
constexpr int OB_EST_QUAT = 86;   // [4] pelvis.orientation (IMU quaternion through its rotation matrix and back: +-q)

// derived-quantity row (optional, cassie_batch_enable_aux): the reference's read-only queries (src/cassiemujoco.c:1586-1961) as by-products
// in-kernel estimator (optional, extended instance): per environment a row of doubles with the filter state (doubles in every precision: the
// covariance recursion is the reference's unsymmetrised one and needs them) and a row of reals with its outputs
constexpr int EST_W = 128;      // doubles: [0] started, [1..42] x axis (state 6, covariance 36), [43..84] y axis, [85..89] z state, [90..114] z covariance,
constexpr int ES_X = 1, ES_Z = 85, ES_PZ = 90, ES_TERRAIN = 115, ES_FORCE = 116;   // [115] terrain, [116..121] leg forces (toeForce of each foot, double)
constexpr int OB_EST_OUT = 96;  // [EO_W] the estimator's force model and filters, columns of the observation row (zero until cassie_batch_enable_estimator_device)
constexpr int EO_W = 16;        // reals: pelvis.position 3, translationalVelocity 3, externalForce 3, terrain.height 1, toeForce (= heelForce) L 3, R 3
constexpr int EO_POS = 0, EO_VEL = 3, EO_EXTF = 6, EO_TERRAIN = 9, EO_TOE = 10;
constexpr int AUX_W = 64;
constexpr int AX_FOOT_FORCE = 0;                    // [12] cassie_sim_foot_forces: left xyz, 3 zeros, right xyz, 3 zeros
constexpr int AX_TOE_FORCE = 12, AX_HEEL_FORCE = 18; // [6] [6] cassie_sim_heeltoe_forces: left xyz, right xyz
constexpr int AX_FOOT_POS = 24;                     // [6]  cassie_sim_foot_positions
constexpr int AX_FOOT_VEL = 30;                     // [12] cassie_sim_foot_velocities
constexpr int AX_CM_POS = 42, AX_CM_VEL = 45, AX_ANGMOM = 48;   // [3] each: centre of mass, its velocity, angular momentum about it
constexpr int AX_OBSTACLE = 51, AX_SELF = 52, AX_GROUPMASK = 53, AX_NCON = 54;
constexpr int AX_TMP = 56;                          // [8] toe / heel world xy of both feet, carried between stages of one sub-step

// per-environment model constants (optional, domain randomisation: src/cassiemujoco.c:1303-1436 setters + mj_setConst :949-977).
// When a batch carries these rows the step kernel reads the listed constants from the env's row instead of the shared model block.
constexpr int CE_W = 288;
constexpr int CE_MASS = 0;       // [32] body_mass
constexpr int CE_IPOS = 32;      // [32][3] body_ipos
constexpr int CE_DAMP = 128;     // [32] dof_damping (main tree)
constexpr int CE_FRIC = 160;     // [32] sliding friction per collision geom (device geom ids: moving 0..15, static 16..31)
constexpr int CE_BINVW = 192;    // [32] body_invweight0 (translational)  -- written by the set_const launch
constexpr int CE_DINVW = 224;    // [32] dof_invweight0                    -- written by the set_const launch
constexpr int CE_ROOT_MINV = 256, CE_TOT_MINV = 257, CE_PGS_SCALE = 258;   // 1 / main-tree mass, 1 / total mass, 1 / (meaninertia * nv)

// ---- per-warp scratch (in units of `real`)
constexpr int S_XPOS = 0;                       // [32][3]
constexpr int S_XQUAT = S_XPOS + 96;            // [32][4]
constexpr int S_XMAT = S_XQUAT + 128;           // [32][9]
constexpr int S_CDOF = S_XMAT + 288;            // [32][6]
constexpr int S_QLD = S_CDOF + 192;             // [320] (qM itself lives in a global scratch row)
constexpr int S_DINV = S_QLD + NM_MAX;          // [32]
constexpr int S_DSQI = S_DINV + 32;             // [32]
constexpr int S_QPOS = S_DSQI + 32;             // [44]
constexpr int S_VEC = S_QPOS + 44;              // [6][32] general vectors ([96..127]: IMU stash; [128..159]: exchange buffer; [160..183]: extra-body qvel / qacc_smooth / qacc_ws / qacc)
constexpr int S_CON = S_VEC + 192;              // [MAXCON][16]
constexpr int S_EFC = S_CON + MAXCON * 16;               // [NEFC][4]: row-build scalars {.., pos, src, ineq} then solver constants {b, 1/A, A, +-R}
constexpr int S_Y = S_EFC + 4 * NEFC;           // [NEFC][ystride] constraint matrix; before the constraint stage it holds the temporaries below
CASSIE_HD inline constexpr int scratch_reals(int ystride) { return S_Y + NEFC * ystride; }
// the extended instance (derived-quantity rows) appends a [32][6] copy of cdof that outlives the constraint stage
CASSIE_HD inline constexpr int scratch_reals_ext(int ystride) { return scratch_reals(ystride) + 192; }
// temporaries inside the S_Y region (dead before the first constraint row is written)
constexpr int T_CINERT = 0;                     // [32][10]
constexpr int T_CRB = 320;                      // [32][10]; during kinematics: xanchor[32][3], xaxis[32][3], qloc[32][4]; later chain sums [32][6]
constexpr int T_CVEL = 640;                     // [32][6]
constexpr int T_CFRC = 832;                     // [32][6]
constexpr int T_CDOFD = 1024;                   // [32][6]
constexpr int T_GEOM = 1216;                    // [16][12] geom world poses (position + rotation by columns), live from the collision stage until the contact list is written
static_assert(T_GEOM + 192 <= NEFC * YSTRIDE_MAIN, "temporaries must fit in the constraint-matrix region");
static_assert(NEFC >= 48 && 16 * YSTRIDE_MAX <= S_QLD - S_XPOS, "the dense solver path keeps A in rows 32..47 of Y and in the kinematics buffers");
// slots of a row's 4 scalars while the rows are being built (overwritten by the solver constants afterwards)
constexpr int E_POS = 0, E_SRC = 1, E_INEQ = 2, E_SIDE = 3;   // E_SIDE: 0 / 1 the row touches the base and at most the first leg, 2 ... the second leg, 3 both legs


// ---- debug dump (tests only; one block per env, in `real`)
constexpr int D_XPOS = 0, D_XQUAT = 96, D_CDOF = 224, D_QM = 416, D_QLD = 736, D_BIAS = 1056, D_PASSIVE = 1088, D_SMOOTH = 1120,
              D_QACCS = 1152, D_QACC = 1184, D_QFRCC = 1216, D_COUNTS = 1248, D_EFC_B = 1252, D_EFC_F = 1316, D_EFC_R = 1380,
              D_EFC_AREF = 1444, D_SENS = 1508, D_J = 1540, D_SIZE = 3600;

}  // namespace cassie
