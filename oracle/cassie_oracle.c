/* cassie_oracle.c -- fp64 CPU ORACLE for the cassie_sim_step_pd hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (cassie-mujoco-sim_b200/, include/) may link,
 * import or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * What it restates (reference = osudrl/cassie-mujoco-sim @ /root/reference):
 *   - glue and ordering of cassie_sim_step_pd        src/cassiemujoco.c:1115-1157
 *   - robot I/O emulation (encoders, filters, motor) src/cassiemujoco.c:194-208, 558-664, 737-803
 *   - init / reset state                             src/cassiemujoco.c:695-734, 979-1036, 2008-2033
 *   - the physics the reference obtains from the un-vendored binary MuJoCo 2.1.0
 *     (mj_step1/mj_step2, src/cassiemujoco.c:1132-1133): kinematics, comPos, CRB, L'DL,
 *     collision (plane-sphere/capsule, capsule-capsule), constraint assembly (connect, joint
 *     limit, pyramidal contact), PGS with warm start, sensors, implicit-damping Euler.
 *     This is a restatement of MuJoCo's published algorithm (Computation chapter + the
 *     open-sourced engine_*.c), specialised to the features the Cassie models use.
 *   - the closed Agility blocks pd_input_step (motor-PD and task-PD branches), cassie_core_sim_step
 *     (safety layer) and state_output_step (estimator: leg kinematics, spring-force model, Kalman
 *     filters), include/pd_input.h:34, include/cassie_core_sim.h:34, include/state_output.h:33-34,
 *     from black-box probing and, for the filters, the block's own memory (SURVEY.md section
 *     8a-2/8a-3/8a-8).  With -DORACLE_USE_AGILITY_REF the real archive src/libagilitycassie.a is
 *     linked instead (oracle/_ref/liboracle_ref.so) and the three blocks are the real ones.
 *
 * PARITY STATUS: "parity unpinned" for the physics -- no MuJoCo of any version is reachable on the
 * build container or on the GPU box (tools/probe_reference.py; committed logs
 * profiles/r2_probe_reference_{build_container,gpu_box}.json), and the reference ships no golden
 * trajectories.  tests/test_mujoco_parity.py diffs this file stage by stage
 * against a real MuJoCo the day one is importable.  The Agility-block twins ARE pinned against the
 * archive (tests/test_agility_twins.py).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/cassie_bus.h"

#define MINVAL 1e-15
#define MAXV 48
#define MAXB 40
#define MAXJ 40
#define MAXG 64
#define MAXCON 64
#define MAXEFC 160
#define MAXNM 512
enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { G_PLANE = 0, G_HFIELD = 1, G_SPHERE = 2, G_CAPSULE = 3, G_BOX = 6 };
enum { C_EQUALITY = 0, C_LIMIT = 3, C_FRICTIONLESS = 5, C_PYRAMIDAL = 6 };

typedef struct {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, neq, nM, nsensor, nhfield;
  double timestep, gravity[3], magnetic[3], tolerance, impratio, meaninertia;
  int iterations;
  int body_parentid[MAXB], body_rootid[MAXB], body_weldid[MAXB], body_jntnum[MAXB], body_jntadr[MAXB],
      body_dofnum[MAXB], body_dofadr[MAXB];
  double body_pos[MAXB][3], body_quat[MAXB][4], body_ipos[MAXB][3], body_iquat[MAXB][4], body_mass[MAXB],
      body_inertia[MAXB][3], body_invweight0[MAXB][2], body_subtreemass[MAXB];
  int jnt_type[MAXJ], jnt_qposadr[MAXJ], jnt_dofadr[MAXJ], jnt_bodyid[MAXJ], jnt_limited[MAXJ];
  double jnt_pos[MAXJ][3], jnt_axis[MAXJ][3], jnt_stiffness[MAXJ], jnt_range[MAXJ][2], jnt_margin[MAXJ],
      jnt_solref[MAXJ][2], jnt_solimp[MAXJ][5];
  int dof_bodyid[MAXV], dof_jntid[MAXV], dof_parentid[MAXV], dof_Madr[MAXV];
  double dof_armature[MAXV], dof_damping[MAXV], dof_invweight0[MAXV];
  double qpos0[MAXV + 8], qpos_spring[MAXV + 8];
  int geom_type[MAXG], geom_bodyid[MAXG], geom_contype[MAXG], geom_conaffinity[MAXG], geom_condim[MAXG],
      geom_priority[MAXG], geom_hfid[MAXG], geom_user[MAXG], geom_group[MAXG];
  double geom_pos[MAXG][3], geom_quat[MAXG][4], geom_size[MAXG][3], geom_friction[MAXG][3], geom_solref[MAXG][2],
      geom_solimp[MAXG][5], geom_rbound[MAXG], geom_solmix[MAXG], geom_margin[MAXG], geom_gap[MAXG];
  int site_bodyid[32];
  double site_pos[32][3], site_quat[32][4];
  int eq_obj1id[8], eq_obj2id[8];
  double eq_data[8][6], eq_solref[8][2], eq_solimp[8][5];
  int actuator_jntid[16], actuator_ctrllimited[16];
  double actuator_gear[16], actuator_ctrlrange[16][2], actuator_user[16];
  int sensor_type[32], sensor_objid[32];
  double sensor_user[32], sensor_cutoff[32];
  int hfield_nrow, hfield_ncol;
  double hfield_size[4];
  float *hfield_data;
  int imu_site;
  /* ids looked up by name at init (src/cassiemujoco.c:861-866); -1 when the model has no such site */
  int left_foot_body, right_foot_body, left_heel, left_toe, right_heel, right_toe;
  double toe_local[2][3], heel_local[2][3]; /* toe / heel points in the foot body frames (sites, else the foot capsule's end points) */
  int ncand, cand_ready; short cand[1024][2]; /* geom pairs that pass the static filters (weld / parent-child / contype-conaffinity), in MuJoCo's order */
} OModel;

typedef struct {
  double pos[3], frame[9], dist, friction[5], solref[2], solimp[5], includemargin, mu;
  int dim, geom1, geom2, efc_address;
} OContact;

typedef struct {
  double time;
  double qpos[MAXV + 8], qvel[MAXV], qacc[MAXV], qacc_warmstart[MAXV], ctrl[16], xfrc_applied[MAXB][6],
      qfrc_applied[MAXV];
  /* position-dependent */
  double xpos[MAXB][3], xquat[MAXB][4], xmat[MAXB][9], xipos[MAXB][3], ximat[MAXB][9], xanchor[MAXJ][3],
      xaxis[MAXJ][3], geom_xpos[MAXG][3], geom_xmat[MAXG][9], site_xpos[32][3], site_xmat[32][9],
      subtree_com[MAXB][3], cdof[MAXV][6], cinert[MAXB][10], crb[MAXB][10], qM[MAXNM], qLD[MAXNM],
      qLDiagInv[MAXV], qLDiagSqrtInv[MAXV];
  int ncon, nefc, ne, nl;
  OContact contact[MAXCON];
  int efc_type[MAXEFC], efc_id[MAXEFC];
  double efc_J[MAXEFC][MAXV], efc_pos[MAXEFC], efc_margin[MAXEFC], efc_diagApprox[MAXEFC], efc_R[MAXEFC],
      efc_D[MAXEFC], efc_KBIP[MAXEFC][4], efc_vel[MAXEFC], efc_aref[MAXEFC], efc_b[MAXEFC], efc_force[MAXEFC];
  double *efc_AR, *efc_JM2; /* nefc x nefc, nefc x nv (allocated once) */
  /* velocity-dependent */
  double cvel[MAXB][6], cdof_dot[MAXV][6], qfrc_bias[MAXV], qfrc_passive[MAXV], actuator_velocity[16],
      actuator_length[16], actuator_force[16], qfrc_actuator[MAXV], qfrc_smooth[MAXV], qacc_smooth[MAXV],
      qfrc_constraint[MAXV], cacc[MAXB][6], subtree_linvel[MAXB][3], subtree_angmom[MAXB][3];
  double sensordata[40];
  int solver_iter, unsupported_pairs, dropped_contacts;
  int cap_con, cap_rows; /* optional: the product's capacities (12 contacts, 48 rows; 0 = MuJoCo-like, unlimited up to MAXCON / MAXEFC), so that overflow
                            situations can be compared too.  Contacts beyond the cap are dropped in contact order; limit rows beyond the row cap
                            are dropped; the first contact whose rows do not fit ends the contact rows. */
} OData;

/* ------------------------------------------------------------------ small vector math */
static void zero(double *x, int n) { memset(x, 0, sizeof(double) * n); }
static void copyv(double *d, const double *s, int n) { memcpy(d, s, sizeof(double) * n); }
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double dotn(const double *a, const double *b, int n) { double s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; }
static void cross(double *r, const double *a, const double *b) {
  r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
}
static double normalize3(double *v) {
  double n = sqrt(dot3(v, v));
  if (n < MINVAL) { v[0] = 1; v[1] = 0; v[2] = 0; return 0; }
  v[0] /= n; v[1] /= n; v[2] /= n; return n;
}
static double normalize4(double *q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return 0; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n; return n;
}
static void mulQuat(double *r, const double *a, const double *b) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  copyv(r, t, 4);
}
static void quat2Mat(double *m, const double *q) {
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3], q11 = q[1] * q[1], q12 = q[1] * q[2],
         q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02); m[3] = 2 * (q12 + q03);
  m[5] = 2 * (q23 - q01); m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
static void rotVecQuat(double *r, const double *v, const double *q) {
  double m[9]; quat2Mat(m, q);
  double t[3] = {m[0] * v[0] + m[1] * v[1] + m[2] * v[2], m[3] * v[0] + m[4] * v[1] + m[5] * v[2], m[6] * v[0] + m[7] * v[1] + m[8] * v[2]};
  copyv(r, t, 3);
}
static void mulMatVec3(double *r, const double *m, const double *v) {
  double t[3] = {m[0] * v[0] + m[1] * v[1] + m[2] * v[2], m[3] * v[0] + m[4] * v[1] + m[5] * v[2], m[6] * v[0] + m[7] * v[1] + m[8] * v[2]};
  copyv(r, t, 3);
}
static void mulMatTVec3(double *r, const double *m, const double *v) {
  double t[3] = {m[0] * v[0] + m[3] * v[1] + m[6] * v[2], m[1] * v[0] + m[4] * v[1] + m[7] * v[2], m[2] * v[0] + m[5] * v[1] + m[8] * v[2]};
  copyv(r, t, 3);
}
static void mulMatMat3(double *r, const double *a, const double *b) {
  double t[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  copyv(r, t, 9);
}
static void axisAngle2Quat(double *q, const double *axis, double ang) {
  if (ang == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  double s = sin(ang * 0.5);
  q[0] = cos(ang * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* spatial algebra in MuJoCo's (angular; linear) layout, 10-number inertias */
static void mulInertVec(double *r, const double *i, const double *v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
static void crossMotion(double *r, const double *vel, const double *v) {
  cross(r, vel, v);
  double t[3]; cross(r + 3, vel, v + 3); cross(t, vel + 3, v);
  r[3] += t[0]; r[4] += t[1]; r[5] += t[2];
}
static void crossForce(double *r, const double *vel, const double *f) {
  double t[3]; cross(r, vel, f); cross(t, vel + 3, f + 3);
  r[0] += t[0]; r[1] += t[1]; r[2] += t[2];
  cross(r + 3, vel, f + 3);
}

/* ------------------------------------------------------------------ model loading */
static int rd_i(char *tok[], int n, int *dst, int cap) { if (n > cap) n = cap; for (int i = 0; i < n; i++) dst[i] = atoi(tok[i]); return n; }
static int rd_f(char *tok[], int n, double *dst, int cap) { if (n > cap) n = cap; for (int i = 0; i < n; i++) dst[i] = strtod(tok[i], NULL); return n; }

OModel *omodel_load(const char *path) {
  FILE *f = fopen(path, "r");
  if (!f) { fprintf(stderr, "oracle: cannot open %s\n", path); return NULL; }
  OModel *m = calloc(1, sizeof(OModel));
  size_t cap = 1 << 20; char *line = malloc(cap);
  static char *tok[70000];
  char site_names[32][64], body_names[MAXB][64]; int nsn = 0, nbn = 0;
  while (fgets(line, cap, f)) {
    if (line[0] == '#') continue;
    char *key = strtok(line, " \n"); if (!key) continue;
    char *ty = strtok(NULL, " \n"); int n = atoi(strtok(NULL, " \n")); (void)ty;
    int nt = 0; char *t;
    while ((t = strtok(NULL, " \n")) && nt < 70000) tok[nt++] = t;
    if (nt < n) n = nt;
#define KI(name, dst, cap_) else if (!strcmp(key, name)) rd_i(tok, n, (int *)(dst), cap_)
#define KF(name, dst, cap_) else if (!strcmp(key, name)) rd_f(tok, n, (double *)(dst), cap_)
    if (0) {}
    KI("nq", &m->nq, 1); KI("nv", &m->nv, 1); KI("nu", &m->nu, 1); KI("nbody", &m->nbody, 1); KI("njnt", &m->njnt, 1);
    KI("ngeom", &m->ngeom, 1); KI("nsite", &m->nsite, 1); KI("neq", &m->neq, 1); KI("nM", &m->nM, 1);
    KI("nsensor", &m->nsensor, 1); KI("nhfield", &m->nhfield, 1);
    KF("opt_timestep", &m->timestep, 1); KF("opt_gravity", m->gravity, 3); KF("opt_magnetic", m->magnetic, 3);
    KF("opt_tolerance", &m->tolerance, 1); KF("opt_impratio", &m->impratio, 1); KI("opt_iterations", &m->iterations, 1);
    KF("stat_meaninertia", &m->meaninertia, 1);
    KI("body_parentid", m->body_parentid, MAXB); KI("body_rootid", m->body_rootid, MAXB); KI("body_weldid", m->body_weldid, MAXB);
    KI("body_jntnum", m->body_jntnum, MAXB); KI("body_jntadr", m->body_jntadr, MAXB); KI("body_dofnum", m->body_dofnum, MAXB);
    KI("body_dofadr", m->body_dofadr, MAXB);
    KF("body_pos", m->body_pos, MAXB * 3); KF("body_quat", m->body_quat, MAXB * 4); KF("body_ipos", m->body_ipos, MAXB * 3);
    KF("body_iquat", m->body_iquat, MAXB * 4); KF("body_mass", m->body_mass, MAXB); KF("body_inertia", m->body_inertia, MAXB * 3);
    KF("body_invweight0", m->body_invweight0, MAXB * 2); KF("body_subtreemass", m->body_subtreemass, MAXB);
    KI("jnt_type", m->jnt_type, MAXJ); KI("jnt_qposadr", m->jnt_qposadr, MAXJ); KI("jnt_dofadr", m->jnt_dofadr, MAXJ);
    KI("jnt_bodyid", m->jnt_bodyid, MAXJ); KI("jnt_limited", m->jnt_limited, MAXJ);
    KF("jnt_pos", m->jnt_pos, MAXJ * 3); KF("jnt_axis", m->jnt_axis, MAXJ * 3); KF("jnt_stiffness", m->jnt_stiffness, MAXJ);
    KF("jnt_range", m->jnt_range, MAXJ * 2); KF("jnt_margin", m->jnt_margin, MAXJ); KF("jnt_solref", m->jnt_solref, MAXJ * 2);
    KF("jnt_solimp", m->jnt_solimp, MAXJ * 5);
    KI("dof_bodyid", m->dof_bodyid, MAXV); KI("dof_jntid", m->dof_jntid, MAXV); KI("dof_parentid", m->dof_parentid, MAXV);
    KI("dof_Madr", m->dof_Madr, MAXV);
    KF("dof_armature", m->dof_armature, MAXV); KF("dof_damping", m->dof_damping, MAXV); KF("dof_invweight0", m->dof_invweight0, MAXV);
    KF("qpos0", m->qpos0, MAXV + 8); KF("qpos_spring", m->qpos_spring, MAXV + 8);
    KI("geom_type", m->geom_type, MAXG); KI("geom_bodyid", m->geom_bodyid, MAXG); KI("geom_contype", m->geom_contype, MAXG);
    KI("geom_conaffinity", m->geom_conaffinity, MAXG); KI("geom_condim", m->geom_condim, MAXG);
    KI("geom_priority", m->geom_priority, MAXG); KI("geom_hfid", m->geom_hfid, MAXG);
    KI("geom_user", m->geom_user, MAXG); KI("geom_group", m->geom_group, MAXG);
    KF("geom_pos", m->geom_pos, MAXG * 3); KF("geom_quat", m->geom_quat, MAXG * 4); KF("geom_size", m->geom_size, MAXG * 3);
    KF("geom_friction", m->geom_friction, MAXG * 3); KF("geom_solref", m->geom_solref, MAXG * 2);
    KF("geom_solimp", m->geom_solimp, MAXG * 5); KF("geom_rbound", m->geom_rbound, MAXG); KF("geom_solmix", m->geom_solmix, MAXG);
    KF("geom_margin", m->geom_margin, MAXG); KF("geom_gap", m->geom_gap, MAXG);
    KI("site_bodyid", m->site_bodyid, 32); KF("site_pos", m->site_pos, 96); KF("site_quat", m->site_quat, 128);
    KI("eq_obj1id", m->eq_obj1id, 8); KI("eq_obj2id", m->eq_obj2id, 8); KF("eq_data", m->eq_data, 48);
    KF("eq_solref", m->eq_solref, 16); KF("eq_solimp", m->eq_solimp, 40);
    KI("actuator_jntid", m->actuator_jntid, 16); KI("actuator_ctrllimited", m->actuator_ctrllimited, 16);
    KF("actuator_gear", m->actuator_gear, 16); KF("actuator_ctrlrange", m->actuator_ctrlrange, 32); KF("actuator_user", m->actuator_user, 16);
    KI("sensor_type", m->sensor_type, 32); KI("sensor_objid", m->sensor_objid, 32);
    KF("sensor_user", m->sensor_user, 32); KF("sensor_cutoff", m->sensor_cutoff, 32);
    KI("hfield_nrow", &m->hfield_nrow, 1); KI("hfield_ncol", &m->hfield_ncol, 1); KF("hfield_size", m->hfield_size, 4);
    else if (!strcmp(key, "names_body")) { for (int i = 0; i < n && i < MAXB; i++) { strncpy(body_names[i], tok[i], 63); body_names[i][63] = 0; } nbn = n < MAXB ? n : MAXB; }
    else if (!strcmp(key, "names_site")) { for (int i = 0; i < n && i < 32; i++) { strncpy(site_names[i], tok[i], 63); site_names[i][63] = 0; } nsn = n < 32 ? n : 32; }
  }
  fclose(f); free(line);
  m->imu_site = 0;
  for (int i = 0; i < nsn; i++) if (!strcmp(site_names[i], "imu")) m->imu_site = i;
  m->left_foot_body = m->right_foot_body = m->left_heel = m->left_toe = m->right_heel = m->right_toe = -1;
  for (int i = 0; i < nbn; i++) { if (!strcmp(body_names[i], "left-foot")) m->left_foot_body = i; if (!strcmp(body_names[i], "right-foot")) m->right_foot_body = i; }
  for (int i = 0; i < nsn; i++) {
    if (!strcmp(site_names[i], "left-heel")) m->left_heel = i; if (!strcmp(site_names[i], "left-toe")) m->left_toe = i;
    if (!strcmp(site_names[i], "right-heel")) m->right_heel = i; if (!strcmp(site_names[i], "right-toe")) m->right_toe = i;
  }
  for (int s = 0; s < 2; s++) { /* toe / heel points: the named sites; models without them (cassie_hfield.xml, cassie_tray_box.xml) fall back
                                   to the end points of the foot capsule, which is where cassie.xml puts the sites (model/cassie.xml:151-154) */
    int fb = s ? m->right_foot_body : m->left_foot_body, toe = s ? m->right_toe : m->left_toe, heel = s ? m->right_heel : m->left_heel;
    if (toe >= 0 && heel >= 0) { for (int k = 0; k < 3; k++) { m->toe_local[s][k] = m->site_pos[toe][k]; m->heel_local[s][k] = m->site_pos[heel][k]; } continue; }
    for (int g = 0; g < m->ngeom; g++) if (m->geom_bodyid[g] == fb && m->geom_type[g] == 3 /* capsule */ && m->geom_contype[g]) {
      const double *q = m->geom_quat[g]; /* z axis of the geom frame */
      double z[3] = {2 * (q[1] * q[3] + q[0] * q[2]), 2 * (q[2] * q[3] - q[0] * q[1]), 1 - 2 * (q[1] * q[1] + q[2] * q[2])};
      for (int k = 0; k < 3; k++) { m->toe_local[s][k] = m->geom_pos[g][k] + z[k] * m->geom_size[g][1]; m->heel_local[s][k] = m->geom_pos[g][k] - z[k] * m->geom_size[g][1]; }
    }
  }
  if (m->nhfield) m->hfield_data = calloc((size_t)m->hfield_nrow * m->hfield_ncol, sizeof(float));
  if (m->nv > MAXV || m->nbody > MAXB || m->ngeom > MAXG || m->nM > MAXNM) { fprintf(stderr, "oracle: model too large\n"); free(m); return NULL; }
  return m;
}
void omodel_free(OModel *m) { if (m) { free(m->hfield_data); free(m); } }

/* ------------------------------------------------------------------ position stage (mj_fwdPosition) */
static void o_kinematics(const OModel *m, OData *d) {
  zero(d->xpos[0], 3); d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  quat2Mat(d->xmat[0], d->xquat[0]); zero(d->xipos[0], 3); copyv(d->ximat[0], d->xmat[0], 9);
  for (int i = 1; i < m->nbody; i++) {
    double xpos[3], xquat[4];
    int jadr = m->body_jntadr[i];
    if (m->body_jntnum[i] == 1 && m->jnt_type[jadr] == JNT_FREE) {
      int qa = m->jnt_qposadr[jadr];
      copyv(xpos, d->qpos + qa, 3); copyv(xquat, d->qpos + qa + 3, 4); normalize4(xquat);
      copyv(d->xanchor[jadr], xpos, 3); copyv(d->xaxis[jadr], m->jnt_axis[jadr], 3);
    } else {
      int pid = m->body_parentid[i];
      if (pid) {
        double v[3]; mulMatVec3(v, d->xmat[pid], m->body_pos[i]);
        for (int k = 0; k < 3; k++) xpos[k] = d->xpos[pid][k] + v[k];
        mulQuat(xquat, d->xquat[pid], m->body_quat[i]);
      } else { copyv(xpos, m->body_pos[i], 3); copyv(xquat, m->body_quat[i], 4); }
      for (int j = 0; j < m->body_jntnum[i]; j++) {
        int jid = jadr + j, qa = m->jnt_qposadr[jid];
        rotVecQuat(d->xaxis[jid], m->jnt_axis[jid], xquat);
        double v[3]; rotVecQuat(v, m->jnt_pos[jid], xquat);
        for (int k = 0; k < 3; k++) d->xanchor[jid][k] = xpos[k] + v[k];
        if (m->jnt_type[jid] == JNT_SLIDE) {
          double s = d->qpos[qa] - m->qpos0[qa];
          for (int k = 0; k < 3; k++) xpos[k] += d->xaxis[jid][k] * s;
        } else {
          double qloc[4];
          if (m->jnt_type[jid] == JNT_BALL) { copyv(qloc, d->qpos + qa, 4); normalize4(qloc); }
          else axisAngle2Quat(qloc, m->jnt_axis[jid], d->qpos[qa] - m->qpos0[qa]);
          mulQuat(xquat, xquat, qloc);
          rotVecQuat(v, m->jnt_pos[jid], xquat);
          for (int k = 0; k < 3; k++) xpos[k] = d->xanchor[jid][k] - v[k];
        }
      }
    }
    normalize4(xquat);
    copyv(d->xpos[i], xpos, 3); copyv(d->xquat[i], xquat, 4); quat2Mat(d->xmat[i], xquat);
    double v[3], iq[4];
    mulMatVec3(v, d->xmat[i], m->body_ipos[i]);
    for (int k = 0; k < 3; k++) d->xipos[i][k] = xpos[k] + v[k];
    mulQuat(iq, xquat, m->body_iquat[i]); quat2Mat(d->ximat[i], iq);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g]; double v[3], q[4];
    mulMatVec3(v, d->xmat[b], m->geom_pos[g]);
    for (int k = 0; k < 3; k++) d->geom_xpos[g][k] = d->xpos[b][k] + v[k];
    mulQuat(q, d->xquat[b], m->geom_quat[g]); quat2Mat(d->geom_xmat[g], q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s]; double v[3], q[4];
    mulMatVec3(v, d->xmat[b], m->site_pos[s]);
    for (int k = 0; k < 3; k++) d->site_xpos[s][k] = d->xpos[b][k] + v[k];
    mulQuat(q, d->xquat[b], m->site_quat[s]); quat2Mat(d->site_xmat[s], q);
  }
}

static void o_comPos(const OModel *m, OData *d) {
  for (int i = 0; i < m->nbody; i++) zero(d->subtree_com[i], 3);
  for (int i = m->nbody - 1; i >= 0; i--) {
    for (int k = 0; k < 3; k++) d->subtree_com[i][k] += d->xipos[i][k] * m->body_mass[i];
    if (i) for (int k = 0; k < 3; k++) d->subtree_com[m->body_parentid[i]][k] += d->subtree_com[i][k];
    if (m->body_subtreemass[i] < MINVAL) copyv(d->subtree_com[i], d->xipos[i], 3);
    else for (int k = 0; k < 3; k++) d->subtree_com[i][k] /= m->body_subtreemass[i];
  }
  zero(d->cinert[0], 10);
  for (int i = 1; i < m->nbody; i++) {
    double off[3]; const double *com = d->subtree_com[m->body_rootid[i]];
    for (int k = 0; k < 3; k++) off[k] = d->xipos[i][k] - com[k];
    const double *R = d->ximat[i], *I = m->body_inertia[i]; double mass = m->body_mass[i], *r = d->cinert[i];
    /* R diag(I) R' in (xx yy zz xy xz yz) order, then parallel axis */
    r[0] = R[0] * R[0] * I[0] + R[1] * R[1] * I[1] + R[2] * R[2] * I[2];
    r[1] = R[3] * R[3] * I[0] + R[4] * R[4] * I[1] + R[5] * R[5] * I[2];
    r[2] = R[6] * R[6] * I[0] + R[7] * R[7] * I[1] + R[8] * R[8] * I[2];
    r[3] = R[0] * R[3] * I[0] + R[1] * R[4] * I[1] + R[2] * R[5] * I[2];
    r[4] = R[0] * R[6] * I[0] + R[1] * R[7] * I[1] + R[2] * R[8] * I[2];
    r[5] = R[3] * R[6] * I[0] + R[4] * R[7] * I[1] + R[5] * R[8] * I[2];
    r[0] += mass * (off[1] * off[1] + off[2] * off[2]); r[1] += mass * (off[0] * off[0] + off[2] * off[2]);
    r[2] += mass * (off[0] * off[0] + off[1] * off[1]);
    r[3] -= mass * off[0] * off[1]; r[4] -= mass * off[0] * off[2]; r[5] -= mass * off[1] * off[2];
    r[6] = mass * off[0]; r[7] = mass * off[1]; r[8] = mass * off[2]; r[9] = mass;
  }
  for (int j = 0; j < m->njnt; j++) {
    int da = m->jnt_dofadr[j], bi = m->jnt_bodyid[j]; double off[3]; const double *com = d->subtree_com[m->body_rootid[bi]];
    for (int k = 0; k < 3; k++) off[k] = com[k] - d->xanchor[j][k];
    switch (m->jnt_type[j]) {
      case JNT_FREE:
        for (int k = 0; k < 3; k++) { zero(d->cdof[da + k], 6); d->cdof[da + k][3 + k] = 1; }
        da += 3; /* fallthrough */
      case JNT_BALL:
        for (int k = 0; k < 3; k++) {
          double ax[3] = {d->xmat[bi][k], d->xmat[bi][k + 3], d->xmat[bi][k + 6]};
          copyv(d->cdof[da + k], ax, 3); cross(d->cdof[da + k] + 3, ax, off);
        }
        break;
      case JNT_SLIDE: zero(d->cdof[da], 3); copyv(d->cdof[da] + 3, d->xaxis[j], 3); break;
      case JNT_HINGE: copyv(d->cdof[da], d->xaxis[j], 3); cross(d->cdof[da] + 3, d->xaxis[j], off); break;
    }
  }
}

static void o_crb(const OModel *m, OData *d) {
  memcpy(d->crb, d->cinert, sizeof(double) * 10 * m->nbody);
  for (int i = m->nbody - 1; i > 0; i--) if (m->body_parentid[i] > 0)
    for (int k = 0; k < 10; k++) d->crb[m->body_parentid[i]][k] += d->crb[i][k];
  zero(d->qM, m->nM);
  for (int i = 0; i < m->nv; i++) {
    int adr = m->dof_Madr[i]; double buf[6];
    d->qM[adr] = m->dof_armature[i];
    mulInertVec(buf, d->crb[m->dof_bodyid[i]], d->cdof[i]);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) d->qM[adr++] += dotn(d->cdof[j], buf, 6);
  }
}

static void factorI(const OModel *m, double *qLD, double *diaginv, double *sqrtinv) {
  int nv = m->nv;
  for (int k = nv - 1; k >= 0; k--) {
    int kk = m->dof_Madr[k], ki = kk + 1;
    for (int i = m->dof_parentid[k]; i >= 0; i = m->dof_parentid[i], ki++) {
      double tmp = qLD[ki] / qLD[kk];
      int cnt = (i < nv - 1 ? m->dof_Madr[i + 1] : m->nM) - m->dof_Madr[i];
      for (int c = 0; c < cnt; c++) qLD[m->dof_Madr[i] + c] -= qLD[ki + c] * tmp;
      qLD[ki] = tmp;
    }
  }
  for (int i = 0; i < nv; i++) { diaginv[i] = 1.0 / qLD[m->dof_Madr[i]]; if (sqrtinv) sqrtinv[i] = 1.0 / sqrt(qLD[m->dof_Madr[i]]); }
}
static void solveLD(const OModel *m, double *x, const double *qLD, const double *diaginv) {
  int nv = m->nv;
  for (int i = nv - 1; i >= 0; i--) if (x[i] != 0) { int a = m->dof_Madr[i] + 1; for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[j] -= qLD[a++] * x[i]; }
  for (int i = 0; i < nv; i++) x[i] *= diaginv[i];
  for (int i = 0; i < nv; i++) { int a = m->dof_Madr[i] + 1; for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[i] -= qLD[a++] * x[j]; }
}
/* x <- sqrt(inv(D)) inv(L') x */
static void solveM2(const OModel *m, const OData *d, double *x) {
  int nv = m->nv;
  for (int i = nv - 1; i >= 0; i--) if (x[i] != 0) { int a = m->dof_Madr[i] + 1; for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[j] -= d->qLD[a++] * x[i]; }
  for (int i = 0; i < nv; i++) x[i] *= d->qLDiagSqrtInv[i];
}

/* jacobian of a world point attached to body (mj_jac) */
static void o_jac(const OModel *m, const OData *d, double *jacp, double *jacr, const double *point, int body) {
  int nv = m->nv; double off[3];
  if (jacp) zero(jacp, 3 * nv); if (jacr) zero(jacr, 3 * nv);
  for (int k = 0; k < 3; k++) off[k] = point[k] - d->subtree_com[m->body_rootid[body]][k];
  while (body && !m->body_dofnum[body]) body = m->body_parentid[body];
  if (!body) return;
  for (int i = m->body_dofadr[body] + m->body_dofnum[body] - 1; i >= 0; i = m->dof_parentid[i]) {
    if (jacr) for (int k = 0; k < 3; k++) jacr[k * nv + i] = d->cdof[i][k];
    if (jacp) { double t[3]; cross(t, d->cdof[i], off); for (int k = 0; k < 3; k++) jacp[k * nv + i] = d->cdof[i][3 + k] + t[k]; }
  }
}

/* ---------------- collision (mj_collision, specialised to the geom types the Cassie models use) */
static void make_frame(double *f) {
  normalize3(f);
  if (sqrt(dot3(f + 3, f + 3)) < 0.5) { f[3] = f[4] = f[5] = 0; if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1; }
  double s = dot3(f, f + 3);
  for (int k = 0; k < 3; k++) f[3 + k] -= f[k] * s;
  normalize3(f + 3); cross(f + 6, f, f + 3);
}
static int raw_plane_sphere(OContact *c, double margin, const double *ppos, const double *pmat, const double *spos, double r) {
  double n[3] = {pmat[2], pmat[5], pmat[8]}, t[3] = {spos[0] - ppos[0], spos[1] - ppos[1], spos[2] - ppos[2]};
  double cdist = dot3(t, n);
  if (cdist > margin + r) return 0;
  c->dist = cdist - r; copyv(c->frame, n, 3); zero(c->frame + 3, 6);
  for (int k = 0; k < 3; k++) c->pos[k] = spos[k] - n[k] * (c->dist * 0.5 + r);
  return 1;
}
static int raw_sphere_sphere(OContact *c, double margin, const double *p1, double r1, const double *p2, double r2) {
  double dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double cdist = sqrt(dot3(dif, dif));
  if (cdist > margin + r1 + r2) return 0;
  c->dist = cdist - r1 - r2;
  if (cdist < MINVAL) { c->frame[0] = 1; c->frame[1] = c->frame[2] = 0; } else for (int k = 0; k < 3; k++) c->frame[k] = dif[k] / cdist;
  zero(c->frame + 3, 6);
  for (int k = 0; k < 3; k++) c->pos[k] = p1[k] + c->frame[k] * (r1 + c->dist * 0.5);
  return 1;
}
static int col_plane_capsule(const OModel *m, const OData *d, OContact *c, int g1, int g2, double margin) {
  const double *mat2 = d->geom_xmat[g2]; double ax[3] = {mat2[2], mat2[5], mat2[8]}, p[3]; double hl = m->geom_size[g2][1];
  for (int k = 0; k < 3; k++) p[k] = d->geom_xpos[g2][k] + ax[k] * hl;
  int n1 = raw_plane_sphere(c, margin, d->geom_xpos[g1], d->geom_xmat[g1], p, m->geom_size[g2][0]);
  for (int k = 0; k < 3; k++) p[k] = d->geom_xpos[g2][k] - ax[k] * hl;
  int n2 = raw_plane_sphere(c + n1, margin, d->geom_xpos[g1], d->geom_xmat[g1], p, m->geom_size[g2][0]);
  for (int i = 0; i < n1 + n2; i++) copyv(c[i].frame + 3, ax, 3);
  return n1 + n2;
}
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
/* closest point on triangle abc to p (Ericson, Real-Time Collision Detection 5.1.5); returns 1 when it lies strictly inside the face */
static int closest_pt_tri(const double *p, const double *a, const double *b, const double *c, double *q) {
  double ab[3], ac[3], ap[3], bp[3], cp[3];
  for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
  double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0 && d2 <= 0) { copyv(q, a, 3); return 0; }
  for (int k = 0; k < 3; k++) bp[k] = p[k] - b[k];
  double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0 && d4 <= d3) { copyv(q, b, 3); return 0; }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); for (int k = 0; k < 3; k++) q[k] = a[k] + v * ab[k]; return 0; }
  for (int k = 0; k < 3; k++) cp[k] = p[k] - c[k];
  double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0 && d5 <= d6) { copyv(q, c, 3); return 0; }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { double w = d2 / (d2 - d6); for (int k = 0; k < 3; k++) q[k] = a[k] + w * ac[k]; return 0; }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { double w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); for (int k = 0; k < 3; k++) q[k] = b[k] + w * (c[k] - b[k]); return 0; }
  double den = 1.0 / (va + vb + vc), v = vb * den, w = vc * den;
  for (int k = 0; k < 3; k++) q[k] = a[k] + ab[k] * v + ac[k] * w;
  return 1;
}
/* Height field vs sphere.  NOT MuJoCo's prism + MPR scheme (that iterative convex solver cannot be restated bit for bit, SURVEY.md hard
 * part 6): the surface is the same triangulation MuJoCo uses (cell diagonal from (r,c+1) to (r+1,c)), every triangle under the sphere's
 * xy box is tested analytically and the single deepest contact is kept.  hfield frame = geom frame, axis aligned. */
static int raw_hfield_sphere(const OModel *m, OContact *c, double margin, const double *hpos, const double *sp, double r) {
  const double sx = m->hfield_size[0], sy = m->hfield_size[1], sz = m->hfield_size[2]; const int nrow = m->hfield_nrow, ncol = m->hfield_ncol;
  double pl[3] = {sp[0] - hpos[0], sp[1] - hpos[1], sp[2] - hpos[2]};
  if (fabs(pl[0]) > sx + r || fabs(pl[1]) > sy + r || pl[2] - r > sz + margin) return 0;
  double dx = 2 * sx / (ncol - 1), dy = 2 * sy / (nrow - 1);
  int c0 = (int)floor((pl[0] - r + sx) / dx), c1 = (int)floor((pl[0] + r + sx) / dx), r0 = (int)floor((pl[1] - r + sy) / dy), r1 = (int)floor((pl[1] + r + sy) / dy);
  c0 = c0 < 0 ? 0 : (c0 > ncol - 2 ? ncol - 2 : c0); c1 = c1 < 0 ? 0 : (c1 > ncol - 2 ? ncol - 2 : c1);
  r0 = r0 < 0 ? 0 : (r0 > nrow - 2 ? nrow - 2 : r0); r1 = r1 < 0 ? 0 : (r1 > nrow - 2 ? nrow - 2 : r1);
  double best = 1e30, bn[3] = {0, 0, 1};
  for (int rr = r0; rr <= r1; rr++) for (int cc = c0; cc <= c1; cc++) {
    double x0 = -sx + cc * dx, y0 = -sy + rr * dy;
    double h00 = m->hfield_data[rr * ncol + cc] * sz, h10 = m->hfield_data[rr * ncol + cc + 1] * sz, h01 = m->hfield_data[(rr + 1) * ncol + cc] * sz, h11 = m->hfield_data[(rr + 1) * ncol + cc + 1] * sz;
    double v00[3] = {x0, y0, h00}, v10[3] = {x0 + dx, y0, h10}, v01[3] = {x0, y0 + dy, h01}, v11[3] = {x0 + dx, y0 + dy, h11};
    const double *tri[2][3] = {{v00, v10, v01}, {v10, v11, v01}};
    for (int t = 0; t < 2; t++) {
      const double *a = tri[t][0], *b = tri[t][1], *cc3 = tri[t][2]; double e1[3], e2[3], n[3], q[3], dist, nrm[3];
      for (int k = 0; k < 3; k++) { e1[k] = b[k] - a[k]; e2[k] = cc3[k] - a[k]; }
      cross(n, e1, e2); normalize3(n);
      if (closest_pt_tri(pl, a, b, cc3, q)) { double d[3] = {pl[0] - a[0], pl[1] - a[1], pl[2] - a[2]}; dist = dot3(n, d) - r; copyv(nrm, n, 3); }
      else {
        double v[3] = {pl[0] - q[0], pl[1] - q[1], pl[2] - q[2]}, dd = sqrt(dot3(v, v));
        if (dot3(v, n) < 0 || dd < 1e-12) continue;
        dist = dd - r; for (int k = 0; k < 3; k++) nrm[k] = v[k] / dd;
      }
      if (dist < margin && dist < best) { best = dist; copyv(bn, nrm, 3); }
    }
  }
  if (best > 1e29) return 0;
  c->dist = best; copyv(c->frame, bn, 3); zero(c->frame + 3, 6);
  for (int k = 0; k < 3; k++) c->pos[k] = sp[k] - bn[k] * (r + best * 0.5);
  return 1;
}
static int col_hfield_capsule(const OModel *m, const OData *d, OContact *c, int g1, int g2, double margin) {
  const double *mat2 = d->geom_xmat[g2]; double ax[3] = {mat2[2], mat2[5], mat2[8]}, p[3]; double hl = m->geom_size[g2][1];
  for (int k = 0; k < 3; k++) p[k] = d->geom_xpos[g2][k] + ax[k] * hl;
  int n1 = raw_hfield_sphere(m, c, margin, d->geom_xpos[g1], p, m->geom_size[g2][0]);
  for (int k = 0; k < 3; k++) p[k] = d->geom_xpos[g2][k] - ax[k] * hl;
  int n2 = raw_hfield_sphere(m, c + n1, margin, d->geom_xpos[g1], p, m->geom_size[g2][0]);
  for (int i = 0; i < n1 + n2; i++) copyv(c[i].frame + 3, ax, 3);
  return n1 + n2;
}
/* ---- box primitives (cassie_tray_box.xml).  These are analytic definitions of our own, NOT restatements of MuJoCo's mjc_PlaneBox /
 * mjc_SphereBox / mjc_CapsuleBox / mjc_BoxBox (the latter two are long special-case routines that cannot be recalled line by line);
 * DESIGN.md states the rules.  Normals point from geom1 to geom2 like every MuJoCo contact. */
static void box_corner(const double *bpos, const double *bmat, const double *s, int i, double *out) {
  double l[3] = {(i & 1) ? s[0] : -s[0], (i & 2) ? s[1] : -s[1], (i & 4) ? s[2] : -s[2]}, w[3];
  mulMatVec3(w, bmat, l); for (int k = 0; k < 3; k++) out[k] = bpos[k] + w[k];
}
/* plane (g1) vs box (g2): every corner below the plane is a contact, at most 4, in corner order */
static int col_plane_box(const OModel *m, const OData *d, OContact *c, int g1, int g2, double margin) {
  const double *pm = d->geom_xmat[g1], *pp = d->geom_xpos[g1]; double n[3] = {pm[2], pm[5], pm[8]}; int cnt = 0;
  for (int i = 0; i < 8 && cnt < 4; i++) {
    double v[3], t[3]; box_corner(d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2], i, v);
    for (int k = 0; k < 3; k++) t[k] = v[k] - pp[k];
    double dist = dot3(t, n);
    if (dist >= margin) continue;
    c[cnt].dist = dist; copyv(c[cnt].frame, n, 3); zero(c[cnt].frame + 3, 6);
    for (int k = 0; k < 3; k++) c[cnt].pos[k] = v[k] - n[k] * dist * 0.5;
    cnt++;
  }
  return cnt;
}
/* sphere (centre sc, radius r) as geom1 vs box as geom2 */
static int raw_sphere_box(OContact *c, double margin, const double *sc, double r, const double *bpos, const double *bmat, const double *s) {
  double t[3] = {sc[0] - bpos[0], sc[1] - bpos[1], sc[2] - bpos[2]}, cl[3], q[3], nl[3], dist; int inside = 1;
  mulMatTVec3(cl, bmat, t);
  for (int k = 0; k < 3; k++) { q[k] = clampd(cl[k], -s[k], s[k]); if (q[k] != cl[k]) inside = 0; }
  if (inside) { /* centre inside the box: leave through the nearest face */
    int ax = 0; double best = 1e30; for (int k = 0; k < 3; k++) { double pen = s[k] - fabs(cl[k]); if (pen < best) { best = pen; ax = k; } }
    nl[0] = nl[1] = nl[2] = 0; nl[ax] = cl[ax] >= 0 ? 1 : -1; q[ax] = nl[ax] * s[ax]; dist = -best - r;
  } else {
    double dv[3] = {cl[0] - q[0], cl[1] - q[1], cl[2] - q[2]}, len = sqrt(dot3(dv, dv));
    if (len - r >= margin) return 0;
    for (int k = 0; k < 3; k++) nl[k] = dv[k] / len;
    dist = len - r;
  }
  if (dist >= margin) return 0;
  double nw[3], qw[3]; mulMatVec3(nw, bmat, nl); mulMatVec3(qw, bmat, q);     /* nw: box -> sphere */
  c->dist = dist; for (int k = 0; k < 3; k++) { c->frame[k] = -nw[k]; c->pos[k] = bpos[k] + qw[k] + nw[k] * dist * 0.5; }
  zero(c->frame + 3, 6);
  return 1;
}
/* capsule (g1) vs box (g2): the point of the capsule axis closest to the box (alternating projections), then sphere vs box there */
static int col_capsule_box(const OModel *m, const OData *d, OContact *c, int g1, int g2, double margin) {
  const double *m1 = d->geom_xmat[g1], *p1 = d->geom_xpos[g1], *bpos = d->geom_xpos[g2], *bmat = d->geom_xmat[g2], *s = m->geom_size[g2];
  double ax[3] = {m1[2], m1[5], m1[8]}, hl = m->geom_size[g1][1], t = 0, p[3], pl[3], q[3], qw[3], tmp[3];
  for (int k = 0; k < 3; k++) tmp[k] = bpos[k] - p1[k];
  t = clampd(dot3(tmp, ax), -hl, hl);
  for (int it = 0; it < 4; it++) {
    for (int k = 0; k < 3; k++) { p[k] = p1[k] + ax[k] * t; tmp[k] = p[k] - bpos[k]; }
    mulMatTVec3(pl, bmat, tmp); for (int k = 0; k < 3; k++) q[k] = clampd(pl[k], -s[k], s[k]);
    mulMatVec3(qw, bmat, q); for (int k = 0; k < 3; k++) tmp[k] = bpos[k] + qw[k] - p1[k];
    t = clampd(dot3(tmp, ax), -hl, hl);
  }
  for (int k = 0; k < 3; k++) p[k] = p1[k] + ax[k] * t;
  int n = raw_sphere_box(c, margin, p, m->geom_size[g1][0], bpos, bmat, s);
  if (n) copyv(c->frame + 3, ax, 3);
  return n;
}
/* box (g1) vs box (g2): corners of g2 inside g1, then corners of g1 inside g2; at most 4 contacts */
static int col_box_box(const OModel *m, const OData *d, OContact *c, int g1, int g2, double margin) {
  int cnt = 0;
  for (int pass = 0; pass < 2 && cnt < 4; pass++) {
    int ga = pass ? g2 : g1, gb = pass ? g1 : g2;    /* corners of gb tested against the volume of ga */
    const double *apos = d->geom_xpos[ga], *amat = d->geom_xmat[ga], *as = m->geom_size[ga];
    for (int i = 0; i < 8 && cnt < 4; i++) {
      double v[3], t[3], vl[3]; box_corner(d->geom_xpos[gb], d->geom_xmat[gb], m->geom_size[gb], i, v);
      for (int k = 0; k < 3; k++) t[k] = v[k] - apos[k];
      mulMatTVec3(vl, amat, t);
      int ax = -1; double best = 1e30, pen;
      for (int k = 0; k < 3; k++) { pen = as[k] - fabs(vl[k]); if (pen <= -margin) { ax = -1; best = -1; break; } if (pen < best) { best = pen; ax = k; } }
      if (ax < 0) continue;
      double nl[3] = {0, 0, 0}, nw[3]; nl[ax] = vl[ax] >= 0 ? 1 : -1; mulMatVec3(nw, amat, nl);   /* outward normal of ga's face */
      double sgn = pass ? -1.0 : 1.0;   /* pass 0: ga = g1, outward normal already points g1 -> g2; pass 1: ga = g2, flip */
      c[cnt].dist = -best; for (int k = 0; k < 3; k++) { c[cnt].frame[k] = sgn * nw[k]; c[cnt].pos[k] = v[k] + nw[k] * best * 0.5; }
      zero(c[cnt].frame + 3, 6); cnt++;
    }
  }
  return cnt;
}
static int col_capsule_capsule(const OModel *m, const OData *d, OContact *c, int g1, int g2, double margin) {
  const double *m1 = d->geom_xmat[g1], *m2 = d->geom_xmat[g2], *p1 = d->geom_xpos[g1], *p2 = d->geom_xpos[g2];
  double a1[3] = {m1[2], m1[5], m1[8]}, a2[3] = {m2[2], m2[5], m2[8]}, dif[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
  double s1 = m->geom_size[g1][1], s2 = m->geom_size[g2][1], r1 = m->geom_size[g1][0], r2 = m->geom_size[g2][0];
  double ma = dot3(a1, a1), mb = -dot3(a1, a2), mc = dot3(a2, a2), u = -dot3(a1, dif), v = dot3(a2, dif), det = ma * mc - mb * mb;
  double v1[3], v2[3];
  if (fabs(det) >= MINVAL) {
    double x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > s1) { x1 = s1; x2 = (v - mb * s1) / mc; } else if (x1 < -s1) { x1 = -s1; x2 = (v + mb * s1) / mc; }
    if (x2 > s2) { x2 = s2; x1 = clampd((u - mb * s2) / ma, -s1, s1); } else if (x2 < -s2) { x2 = -s2; x1 = clampd((u + mb * s2) / ma, -s1, s1); }
    for (int k = 0; k < 3; k++) { v1[k] = p1[k] + a1[k] * x1; v2[k] = p2[k] + a2[k] * x2; }
    return raw_sphere_sphere(c, margin, v1, r1, v2, r2);
  }
  /* parallel axes: the segment ends of 1 against 2 (at most 2 contacts) */
  int n = 0;
  for (int e = 0; e < 2 && n < 2; e++) {
    double x1 = e ? -s1 : s1, x2 = clampd((v - mb * x1) / mc, -s2, s2);
    for (int k = 0; k < 3; k++) { v1[k] = p1[k] + a1[k] * x1; v2[k] = p2[k] + a2[k] * x2; }
    n += raw_sphere_sphere(c + n, margin, v1, r1, v2, r2);
  }
  return n;
}

static void collide_geoms(const OModel *m, OData *d, int g1, int g2) {
  if (m->geom_type[g1] > m->geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
  if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) return;
  double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]), gap = fmax(m->geom_gap[g1], m->geom_gap[g2]);
  int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
  /* bounding-sphere filter */
  if (m->geom_rbound[g1] > 0 && m->geom_rbound[g2] > 0) {
    double dif[3]; for (int k = 0; k < 3; k++) dif[k] = d->geom_xpos[g1][k] - d->geom_xpos[g2][k];
    double bound = m->geom_rbound[g1] + m->geom_rbound[g2] + margin;
    if (dot3(dif, dif) > bound * bound) return;
  } else if (t1 == G_PLANE && m->geom_rbound[g2] > 0) {
    const double *pm = d->geom_xmat[g1]; double n[3] = {pm[2], pm[5], pm[8]}, dif[3];
    for (int k = 0; k < 3; k++) dif[k] = d->geom_xpos[g2][k] - d->geom_xpos[g1][k];
    if (dot3(dif, n) > m->geom_rbound[g2] + margin) return;
  }
  OContact con[8]; int num = 0;
  if (t1 == G_PLANE && t2 == G_SPHERE) num = raw_plane_sphere(con, margin, d->geom_xpos[g1], d->geom_xmat[g1], d->geom_xpos[g2], m->geom_size[g2][0]);
  else if (t1 == G_PLANE && t2 == G_CAPSULE) num = col_plane_capsule(m, d, con, g1, g2, margin);
  else if (t1 == G_CAPSULE && t2 == G_CAPSULE) num = col_capsule_capsule(m, d, con, g1, g2, margin);
  else if (t1 == G_PLANE && t2 == G_BOX) num = col_plane_box(m, d, con, g1, g2, margin);
  else if (t1 == G_SPHERE && t2 == G_BOX) num = raw_sphere_box(con, margin, d->geom_xpos[g1], m->geom_size[g1][0], d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2]);
  else if (t1 == G_CAPSULE && t2 == G_BOX) num = col_capsule_box(m, d, con, g1, g2, margin);
  else if (t1 == G_BOX && t2 == G_BOX) num = col_box_box(m, d, con, g1, g2, margin);
  else if (t1 == G_HFIELD && t2 == G_SPHERE && m->hfield_data) num = raw_hfield_sphere(m, con, margin, d->geom_xpos[g1], d->geom_xpos[g2], m->geom_size[g2][0]);
  else if (t1 == G_HFIELD && t2 == G_CAPSULE && m->hfield_data) num = col_hfield_capsule(m, d, con, g1, g2, margin);
  else { d->unsupported_pairs++; return; }
  if (!num) return;
  /* contact parameter mixing (mj_contactParam) */
  int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2], dim; double fr[3], solref[2], solimp[5];
  if (p1 != p2) {
    int g = p1 > p2 ? g1 : g2; dim = m->geom_condim[g];
    copyv(fr, m->geom_friction[g], 3); copyv(solref, m->geom_solref[g], 2); copyv(solimp, m->geom_solimp[g], 5);
  } else {
    dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
    double mix, s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2];
    if (s1 >= MINVAL && s2 >= MINVAL) mix = s1 / (s1 + s2); else if (s1 < MINVAL && s2 < MINVAL) mix = 0.5; else mix = s1 < MINVAL ? 0.0 : 1.0;
    if (m->geom_solref[g1][0] > 0 && m->geom_solref[g2][0] > 0) for (int k = 0; k < 2; k++) solref[k] = mix * m->geom_solref[g1][k] + (1 - mix) * m->geom_solref[g2][k];
    else for (int k = 0; k < 2; k++) solref[k] = fmin(m->geom_solref[g1][k], m->geom_solref[g2][k]);
    for (int k = 0; k < 5; k++) solimp[k] = mix * m->geom_solimp[g1][k] + (1 - mix) * m->geom_solimp[g2][k];
    for (int k = 0; k < 3; k++) fr[k] = fmax(m->geom_friction[g1][k], m->geom_friction[g2][k]);
  }
  for (int i = 0; i < num; i++) {
    if (con[i].dist >= margin) continue; /* only penetrating (dist < margin) contacts are kept */
    if (d->ncon >= MAXCON || (d->cap_con && d->ncon >= d->cap_con)) { d->dropped_contacts++; continue; }
    OContact *c = &d->contact[d->ncon++];
    *c = con[i]; c->geom1 = g1; c->geom2 = g2; c->dim = dim; c->includemargin = margin - gap;
    c->friction[0] = c->friction[1] = fr[0]; c->friction[2] = fr[1]; c->friction[3] = c->friction[4] = fr[2];
    copyv(c->solref, solref, 2); copyv(c->solimp, solimp, 5);
    make_frame(c->frame);
  }
}

static void o_collision(const OModel *m_, OData *d) {
  OModel *m = (OModel *)m_;
  d->ncon = 0;
  /* body pairs in ascending (b1,b2) order == MuJoCo's sorted broadphase output; the AABB sweep itself only prunes pairs that the
     bounding-sphere / narrow phase would reject anyway.  The pairs that survive the STATIC filters are listed once per model. */
  if (!m->cand_ready) {
    m->ncand = 0;
    for (int b1 = 0; b1 < m->nbody; b1++) for (int b2 = b1 + 1; b2 < m->nbody; b2++) {
      int w1 = m->body_weldid[b1], w2 = m->body_weldid[b2];
      if (w1 == w2) continue;
      if (w1 && w2 && (m->body_weldid[m->body_parentid[w1]] == w2 || m->body_weldid[m->body_parentid[w2]] == w1)) continue;
      for (int g1 = 0; g1 < m->ngeom; g1++) if (m->geom_bodyid[g1] == b1)
        for (int g2 = 0; g2 < m->ngeom; g2++) if (m->geom_bodyid[g2] == b2) {
          if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) continue;
          if (m->ncand < 1024) { m->cand[m->ncand][0] = (short)g1; m->cand[m->ncand][1] = (short)g2; m->ncand++; }
        }
    }
    m->cand_ready = 1;
  }
  for (int i = 0; i < m->ncand; i++) collide_geoms(m, d, m->cand[i][0], m->cand[i][1]);
}
static int add_row(OData *d, const double *J, int nv, double pos, double margin, int type, int id) {
  if (d->nefc >= MAXEFC) return -1;
  int r = d->nefc++;
  copyv(d->efc_J[r], J, nv); d->efc_pos[r] = pos; d->efc_margin[r] = margin; d->efc_type[r] = type; d->efc_id[r] = id;
  return r;
}
static void get_impedance(const double *solimp, double pos, double margin, double *imp) {
  if (solimp[0] == solimp[1] || solimp[2] <= MINVAL) { *imp = 0.5 * (solimp[0] + solimp[1]); return; }
  double x = fabs((pos - margin) / solimp[2]);
  if (x >= 1) { *imp = solimp[1]; return; }
  if (x <= 0) { *imp = solimp[0]; return; }
  double y, p = solimp[4], mid = solimp[3];
  if (p == 1) y = x;
  else if (x <= mid) y = pow(x, p) / pow(mid, p - 1);
  else y = 1 - pow(1 - x, p) / pow(1 - mid, p - 1);
  *imp = solimp[0] + y * (solimp[1] - solimp[0]);
}
static void o_makeConstraint(const OModel *m, OData *d) {
  int nv = m->nv; double jp1[3 * MAXV], jp2[3 * MAXV], J[4][MAXV];
  d->nefc = d->ne = d->nl = 0;
  /* equality: connect */
  for (int e = 0; e < m->neq; e++) {
    int b1 = m->eq_obj1id[e], b2 = m->eq_obj2id[e]; double p1[3], p2[3], v[3];
    mulMatVec3(v, d->xmat[b1], m->eq_data[e]); for (int k = 0; k < 3; k++) p1[k] = d->xpos[b1][k] + v[k];
    mulMatVec3(v, d->xmat[b2], m->eq_data[e] + 3); for (int k = 0; k < 3; k++) p2[k] = d->xpos[b2][k] + v[k];
    o_jac(m, d, jp1, NULL, p1, b1); o_jac(m, d, jp2, NULL, p2, b2);
    for (int k = 0; k < 3; k++) {
      for (int i = 0; i < nv; i++) J[0][i] = jp1[k * nv + i] - jp2[k * nv + i];
      add_row(d, J[0], nv, p1[k] - p2[k], 0, C_EQUALITY, e);
    }
    d->ne += 3;
  }
  /* joint limits */
  for (int j = 0; j < m->njnt; j++) if (m->jnt_limited[j] && (m->jnt_type[j] == JNT_HINGE || m->jnt_type[j] == JNT_SLIDE)) {
    double value = d->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[j][(side + 1) / 2] - value);
      if (dist < margin && !(d->cap_rows && d->nefc >= d->cap_rows)) { zero(J[0], nv); J[0][m->jnt_dofadr[j]] = -(double)side; add_row(d, J[0], nv, dist, margin, C_LIMIT, j); d->nl++; }
    }
  }
  /* contacts */
  for (int c = 0; c < d->ncon; c++) {
    OContact *con = &d->contact[c]; int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
    if (d->cap_rows && d->nefc + (con->dim > 1 ? 2 * (con->dim - 1) : 1) > d->cap_rows) { d->dropped_contacts += d->ncon - c; d->ncon = c; break; }
    con->efc_address = d->nefc;
    o_jac(m, d, jp1, NULL, con->pos, b1); o_jac(m, d, jp2, NULL, con->pos, b2);
    int nr = con->dim > 1 ? 3 : 1;
    for (int r = 0; r < nr; r++) for (int i = 0; i < nv; i++) {
      double s = 0; for (int k = 0; k < 3; k++) s += con->frame[3 * r + k] * (jp2[k * nv + i] - jp1[k * nv + i]);
      J[r][i] = s;
    }
    if (con->dim == 1) add_row(d, J[0], nv, con->dist, con->includemargin, C_FRICTIONLESS, c);
    else for (int k = 1; k < con->dim; k++) {
      double row[MAXV];
      for (int i = 0; i < nv; i++) row[i] = J[0][i] + con->friction[k - 1] * J[k][i];
      add_row(d, row, nv, con->dist, con->includemargin, C_PYRAMIDAL, c);
      for (int i = 0; i < nv; i++) row[i] = J[0][i] - con->friction[k - 1] * J[k][i];
      add_row(d, row, nv, con->dist, con->includemargin, C_PYRAMIDAL, c);
    }
  }
  /* diagApprox, impedance, R, KBIP (mj_makeImpedance) */
  for (int i = 0; i < d->nefc; i++) {
    const double *solref, *solimp; int id = d->efc_id[i];
    switch (d->efc_type[i]) {
      case C_EQUALITY: d->efc_diagApprox[i] = m->body_invweight0[m->eq_obj1id[id]][0] + m->body_invweight0[m->eq_obj2id[id]][0];
        solref = m->eq_solref[id]; solimp = m->eq_solimp[id]; break;
      case C_LIMIT: d->efc_diagApprox[i] = m->dof_invweight0[m->jnt_dofadr[id]]; solref = m->jnt_solref[id]; solimp = m->jnt_solimp[id]; break;
      default: {
        OContact *con = &d->contact[id];
        double tran = m->body_invweight0[m->geom_bodyid[con->geom1]][0] + m->body_invweight0[m->geom_bodyid[con->geom2]][0];
        if (d->efc_type[i] == C_FRICTIONLESS) d->efc_diagApprox[i] = tran;
        else { /* which of the 2*(dim-1) pyramid rows is this? rows of one contact are contiguous */
          int first = i; while (first > 0 && d->efc_type[first - 1] == C_PYRAMIDAL && d->efc_id[first - 1] == id) first--;
          double fri = con->friction[(i - first) / 2];
          d->efc_diagApprox[i] = tran + fri * fri * tran;
        }
        solref = con->solref; solimp = con->solimp;
      }
    }
    double sr0 = solref[0], sr1 = solref[1];
    if (sr0 > 0) sr0 = fmax(sr0, 2 * m->timestep); /* refsafe */
    double imp; get_impedance(solimp, d->efc_pos[i], d->efc_margin[i], &imp);
    d->efc_R[i] = fmax(MINVAL, (1 - imp) * d->efc_diagApprox[i] / imp);
    if (sr0 > 0) { d->efc_KBIP[i][0] = 1 / fmax(MINVAL, solimp[1] * solimp[1] * sr0 * sr0 * sr1 * sr1); d->efc_KBIP[i][1] = 2 / fmax(MINVAL, solimp[1] * sr0); }
    else { d->efc_KBIP[i][0] = -sr0 / fmax(MINVAL, solimp[1] * solimp[1]); d->efc_KBIP[i][1] = -sr1 / fmax(MINVAL, solimp[1]); }
    d->efc_KBIP[i][2] = imp; d->efc_KBIP[i][3] = 0;
  }
  /* pyramidal contacts: all rows of a contact share R = 2 mu^2 R[first] */
  for (int i = 0; i < d->nefc; i++) if (d->efc_type[i] == C_PYRAMIDAL) {
    OContact *con = &d->contact[d->efc_id[i]]; int n = 2 * (con->dim - 1);
    con->mu = con->friction[0] / sqrt(m->impratio);
    double Rpy = 2 * con->mu * con->mu * d->efc_R[i];
    for (int j = 0; j < n; j++) d->efc_R[i + j] = Rpy;
    i += n - 1;
  }
  for (int i = 0; i < d->nefc; i++) d->efc_D[i] = 1 / d->efc_R[i];
}
static void o_projectConstraint(const OModel *m, OData *d) {
  int n = d->nefc, nv = m->nv;
  if (!n) return;
  if (!d->efc_AR) { d->efc_AR = malloc(sizeof(double) * MAXEFC * MAXEFC); d->efc_JM2 = malloc(sizeof(double) * MAXEFC * MAXV); }
  double *JM2 = d->efc_JM2;
  for (int i = 0; i < n; i++) { copyv(JM2 + i * nv, d->efc_J[i], nv); solveM2(m, d, JM2 + i * nv); }
  for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) { double s = dotn(JM2 + i * nv, JM2 + j * nv, nv); d->efc_AR[i * n + j] = d->efc_AR[j * n + i] = s; }
  for (int i = 0; i < n; i++) d->efc_AR[i * n + i] += d->efc_R[i];
}

/* ---------------- sensors: the 29-number Cassie layout (model/cassie.xml:272-292) */
static void o_sensor(const OModel *m, OData *d, int stage) {
  int adr = 0;
  for (int s = 0; s < m->nsensor; s++) {
    int t = m->sensor_type[s], obj = m->sensor_objid[s], dim = (t <= 1) ? 1 : (t == 2 ? 4 : 3);
    double *out = d->sensordata + adr; int st = (t == 3) ? 1 : (t == 4 ? 2 : 0);
    if (st == stage) {
      if (t == 0) out[0] = d->actuator_length[obj];
      else if (t == 1) out[0] = d->qpos[m->jnt_qposadr[obj]];
      else if (t == 2) { mulQuat(out, d->xquat[m->site_bodyid[obj]], m->site_quat[obj]); }
      else if (t == 5) mulMatTVec3(out, d->site_xmat[obj], m->magnetic);
      else {
        int b = m->site_bodyid[obj]; const double *com = d->subtree_com[m->body_rootid[b]]; double dif[3], vel[6], t3[3];
        for (int k = 0; k < 3; k++) dif[k] = d->site_xpos[obj][k] - com[k];
        copyv(vel, d->cvel[b], 6); cross(t3, dif, d->cvel[b]); for (int k = 0; k < 3; k++) vel[3 + k] -= t3[k];
        if (t == 3) mulMatTVec3(out, d->site_xmat[obj], vel);
        else {
          double acc[6], la[3], lv[3], lw[3], corr[3];
          copyv(acc, d->cacc[b], 6); cross(t3, dif, d->cacc[b]); for (int k = 0; k < 3; k++) acc[3 + k] -= t3[k];
          mulMatTVec3(la, d->site_xmat[obj], acc + 3); mulMatTVec3(lw, d->site_xmat[obj], vel); mulMatTVec3(lv, d->site_xmat[obj], vel + 3);
          cross(corr, lw, lv);
          for (int k = 0; k < 3; k++) out[k] = la[k] + corr[k];
        }
      }
      if (m->sensor_cutoff[s] > 0) for (int k = 0; k < dim; k++) out[k] = clampd(out[k], -m->sensor_cutoff[s], m->sensor_cutoff[s]);
    }
    adr += dim;
  }
}

/* ---------------- velocity stage */
static void o_comVel(const OModel *m, OData *d) {
  zero(d->cvel[0], 6);
  for (int i = 1; i < m->nbody; i++) {
    double cvel[6]; copyv(cvel, d->cvel[m->body_parentid[i]], 6);
    int bda = m->body_dofadr[i];
    for (int j = 0; j < m->body_dofnum[i]; j++) {
      int jt = m->jnt_type[m->dof_jntid[bda + j]];
      if (jt == JNT_FREE) {
        for (int k = 0; k < 3; k++) { zero(d->cdof_dot[bda + k], 6); for (int c = 0; c < 6; c++) cvel[c] += d->cdof[bda + k][c] * d->qvel[bda + k]; }
        j += 3; jt = JNT_BALL;
      }
      if (jt == JNT_BALL) {
        for (int k = 0; k < 3; k++) crossMotion(d->cdof_dot[bda + j + k], cvel, d->cdof[bda + j + k]);
        for (int k = 0; k < 3; k++) for (int c = 0; c < 6; c++) cvel[c] += d->cdof[bda + j + k][c] * d->qvel[bda + j + k];
        j += 2;
      } else {
        crossMotion(d->cdof_dot[bda + j], cvel, d->cdof[bda + j]);
        for (int c = 0; c < 6; c++) cvel[c] += d->cdof[bda + j][c] * d->qvel[bda + j];
      }
    }
    copyv(d->cvel[i], cvel, 6);
  }
}
static void o_passive(const OModel *m, OData *d) {
  zero(d->qfrc_passive, m->nv);
  for (int j = 0; j < m->njnt; j++) if (m->jnt_stiffness[j] != 0 && (m->jnt_type[j] == JNT_HINGE || m->jnt_type[j] == JNT_SLIDE))
    d->qfrc_passive[m->jnt_dofadr[j]] = -m->jnt_stiffness[j] * (d->qpos[m->jnt_qposadr[j]] - m->qpos_spring[m->jnt_qposadr[j]]);
  for (int i = 0; i < m->nv; i++) d->qfrc_passive[i] -= m->dof_damping[i] * d->qvel[i];
}
static void o_rne_bias(const OModel *m, OData *d) {
  double cacc[MAXB][6], cfrc[MAXB][6];
  zero(cacc[0], 3); for (int k = 0; k < 3; k++) cacc[0][3 + k] = -m->gravity[k];
  zero(cfrc[0], 6);
  for (int i = 1; i < m->nbody; i++) {
    int bda = m->body_dofadr[i]; double t[6], t1[6];
    copyv(cacc[i], cacc[m->body_parentid[i]], 6);
    for (int j = 0; j < m->body_dofnum[i]; j++) for (int c = 0; c < 6; c++) cacc[i][c] += d->cdof_dot[bda + j][c] * d->qvel[bda + j];
    mulInertVec(cfrc[i], d->cinert[i], cacc[i]); mulInertVec(t, d->cinert[i], d->cvel[i]); crossForce(t1, d->cvel[i], t);
    for (int c = 0; c < 6; c++) cfrc[i][c] += t1[c];
  }
  for (int i = m->nbody - 1; i > 0; i--) if (m->body_parentid[i]) for (int c = 0; c < 6; c++) cfrc[m->body_parentid[i]][c] += cfrc[i][c];
  for (int i = 0; i < m->nv; i++) d->qfrc_bias[i] = dotn(d->cdof[i], cfrc[m->dof_bodyid[i]], 6);
}

/* ---------------- acceleration stage */
static void o_fwdActuation(const OModel *m, OData *d) {
  zero(d->qfrc_actuator, m->nv);
  for (int i = 0; i < m->nu; i++) {
    double c = d->ctrl[i];
    if (m->actuator_ctrllimited[i]) c = clampd(c, m->actuator_ctrlrange[i][0], m->actuator_ctrlrange[i][1]);
    d->actuator_force[i] = c;
    d->qfrc_actuator[m->jnt_dofadr[m->actuator_jntid[i]]] += m->actuator_gear[i] * c;
  }
}
static void o_fwdAcceleration(const OModel *m, OData *d) {
  int nv = m->nv; double jp[3 * MAXV], jr[3 * MAXV];
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_applied[i] + d->qfrc_actuator[i];
  for (int b = 1; b < m->nbody; b++) {
    const double *x = d->xfrc_applied[b]; int nz = 0; for (int k = 0; k < 6; k++) if (x[k] != 0) nz = 1;
    if (!nz) continue;
    o_jac(m, d, jp, jr, d->xipos[b], b);
    for (int i = 0; i < nv; i++) for (int k = 0; k < 3; k++) d->qfrc_smooth[i] += jp[k * nv + i] * x[k] + jr[k * nv + i] * x[3 + k];
  }
  copyv(d->qacc_smooth, d->qfrc_smooth, nv); solveLD(m, d->qacc_smooth, d->qLD, d->qLDiagInv);
}
static void mulJacVec(const OModel *m, const OData *d, double *res, const double *v) { for (int i = 0; i < d->nefc; i++) res[i] = dotn(d->efc_J[i], v, m->nv); }
static int is_ineq(int t) { return t == C_LIMIT || t == C_FRICTIONLESS || t == C_PYRAMIDAL; }
static void o_fwdConstraint(const OModel *m, OData *d) {
  int nv = m->nv, n = d->nefc; d->solver_iter = 0;
  if (!n) { copyv(d->qacc, d->qacc_smooth, nv); zero(d->qfrc_constraint, nv); return; }
  /* reference acceleration and b */
  mulJacVec(m, d, d->efc_vel, d->qvel);
  for (int i = 0; i < n; i++) d->efc_aref[i] = -d->efc_KBIP[i][1] * d->efc_vel[i] - d->efc_KBIP[i][0] * d->efc_KBIP[i][2] * (d->efc_pos[i] - d->efc_margin[i]);
  mulJacVec(m, d, d->efc_b, d->qacc_smooth);
  for (int i = 0; i < n; i++) d->efc_b[i] -= d->efc_aref[i];
  /* warm start: forces implied by qacc_warmstart, kept only if their dual cost is negative */
  double *f = d->efc_force, *AR = d->efc_AR; double jar[MAXEFC];
  mulJacVec(m, d, jar, d->qacc_warmstart);
  for (int i = 0; i < n; i++) { jar[i] -= d->efc_aref[i]; f[i] = -d->efc_D[i] * jar[i]; if (is_ineq(d->efc_type[i]) && jar[i] >= 0) f[i] = 0; }
  double cost = dotn(f, d->efc_b, n);
  for (int i = 0; i < n; i++) cost += 0.5 * f[i] * dotn(AR + i * n, f, n);
  if (cost > 0) zero(f, n);
  /* PGS */
  double scale = 1 / (m->meaninertia * (nv > 1 ? nv : 1));
  int iter = 0;
  while (iter < m->iterations) {
    double improvement = 0;
    for (int i = 0; i < n; i++) {
      double res = d->efc_b[i] + dotn(AR + i * n, f, n), old = f[i];
      f[i] -= res / AR[i * n + i];
      if (is_ineq(d->efc_type[i]) && f[i] < 0) f[i] = 0;
      double delta = f[i] - old, change = 0.5 * delta * delta * AR[i * n + i] + delta * res;
      if (change > 1e-10) { f[i] = old; change = 0; }
      improvement -= change;
    }
    improvement *= scale; iter++;
    if (improvement < m->tolerance) break;
  }
  d->solver_iter = iter;
  for (int j = 0; j < nv; j++) { double s = 0; for (int i = 0; i < n; i++) s += d->efc_J[i][j] * f[i]; d->qfrc_constraint[j] = s; }
  copyv(d->qacc, d->qfrc_constraint, nv); solveLD(m, d->qacc, d->qLD, d->qLDiagInv);
  for (int j = 0; j < nv; j++) d->qacc[j] += d->qacc_smooth[j];
}
static void o_cacc(const OModel *m, OData *d) { /* the part of mj_rnePostConstraint the accelerometer needs */
  zero(d->cacc[0], 3); for (int k = 0; k < 3; k++) d->cacc[0][3 + k] = -m->gravity[k];
  for (int i = 1; i < m->nbody; i++) {
    int bda = m->body_dofadr[i]; copyv(d->cacc[i], d->cacc[m->body_parentid[i]], 6);
    for (int j = 0; j < m->body_dofnum[i]; j++) for (int c = 0; c < 6; c++) d->cacc[i][c] += d->cdof_dot[bda + j][c] * d->qvel[bda + j] + d->cdof[bda + j][c] * d->qacc[bda + j];
  }
}
static void quatIntegrate(double *q, const double *w, double h) {
  double ax[3] = {w[0], w[1], w[2]}, ang = h * normalize3(ax), qr[4];
  axisAngle2Quat(qr, ax, ang); normalize4(q); mulQuat(q, q, qr);
}
static void o_euler(const OModel *m, OData *d) {
  int nv = m->nv; double MhB[MAXNM], dinv[MAXV], qacc[MAXV];
  copyv(MhB, d->qM, m->nM);
  for (int i = 0; i < nv; i++) MhB[m->dof_Madr[i]] += m->timestep * m->dof_damping[i];
  factorI(m, MhB, dinv, NULL);
  for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
  solveLD(m, qacc, MhB, dinv);
  for (int i = 0; i < nv; i++) d->qvel[i] += m->timestep * qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    switch (m->jnt_type[j]) {
      case JNT_FREE: for (int k = 0; k < 3; k++) d->qpos[qa + k] += m->timestep * d->qvel[da + k]; quatIntegrate(d->qpos + qa + 3, d->qvel + da + 3, m->timestep); break;
      case JNT_BALL: quatIntegrate(d->qpos + qa, d->qvel + da, m->timestep); break;
      default: d->qpos[qa] += m->timestep * d->qvel[da];
    }
  }
  d->time += m->timestep;
  copyv(d->qacc_warmstart, d->qacc, nv);
}

void o_forward(const OModel *m, OData *d) {
  o_kinematics(m, d); o_comPos(m, d); o_crb(m, d);
  copyv(d->qLD, d->qM, m->nM); factorI(m, d->qLD, d->qLDiagInv, d->qLDiagSqrtInv);
  o_collision(m, d); o_makeConstraint(m, d); o_projectConstraint(m, d);
  for (int i = 0; i < m->nu; i++) { int j = m->actuator_jntid[i]; d->actuator_length[i] = m->actuator_gear[i] * d->qpos[m->jnt_qposadr[j]]; }
  o_sensor(m, d, 0);
  for (int i = 0; i < m->nu; i++) d->actuator_velocity[i] = m->actuator_gear[i] * d->qvel[m->jnt_dofadr[m->actuator_jntid[i]]];
  o_comVel(m, d); o_passive(m, d); o_rne_bias(m, d);
  o_sensor(m, d, 1);
  o_fwdActuation(m, d); o_fwdAcceleration(m, d); o_fwdConstraint(m, d);
  o_cacc(m, d); o_sensor(m, d, 2);
}
void o_step(const OModel *m, OData *d) { o_forward(m, d); o_euler(m, d); }

/* ================================================================== Cassie glue (src/cassiemujoco.c) */
#define NUM_DRIVES 10
#define NUM_JOINTS 6
#define DELAY 6
static const int drive_b[9] = {2727, 534, -2658, -795, 72, 110, 19, -6, -3};       /* :198-200 */
static const double joint_b[4] = {12.348, 12.348, -12.348, -12.348};                /* :202-204 */
static const double joint_a[3] = {1.0, -1.7658, 0.79045};                           /* :206-208 */

#ifdef ORACLE_USE_AGILITY_REF
typedef struct CassieCoreSim cassie_core_sim_t; typedef struct StateOutput state_output_t; typedef struct PdInput pd_input_t;
cassie_core_sim_t *cassie_core_sim_alloc(void); void cassie_core_sim_setup(cassie_core_sim_t *); void cassie_core_sim_free(cassie_core_sim_t *);
void cassie_core_sim_step(cassie_core_sim_t *, const cassie_user_in_t *, const cassie_out_t *, cassie_in_t *);
state_output_t *state_output_alloc(void); void state_output_setup(state_output_t *); void state_output_free(state_output_t *);
void state_output_step(state_output_t *, const cassie_out_t *, state_out_t *);
pd_input_t *pd_input_alloc(void); void pd_input_setup(pd_input_t *); void pd_input_free(pd_input_t *);
void pd_input_step(pd_input_t *, const pd_in_t *, const cassie_out_t *, cassie_user_in_t *);
#endif

/* state of the estimator's filters (o_est_filter_step) */
typedef struct { int started; double x[2][6], P[2][36], z[5], Pz[25], terrain; } OEst;
typedef struct {
  OModel *m; OData *d;
  cassie_out_t cassie_out;
  int drive_filter[NUM_DRIVES][9];
  double joint_filter_x[NUM_JOINTS][4], joint_filter_y[NUM_JOINTS][3];
  double torque_delay[NUM_DRIVES][DELAY];
  OEst est2;
#ifdef ORACLE_USE_AGILITY_REF
  cassie_core_sim_t *core; state_output_t *est; pd_input_t *pd;
#endif
} OSim;

static elmo_out_t *drive_ptr(cassie_out_t *o, int i) {
  cassie_leg_out_t *leg = i < 5 ? &o->leftLeg : &o->rightLeg;
  elmo_out_t *t[5] = {&leg->hipRollDrive, &leg->hipYawDrive, &leg->hipPitchDrive, &leg->kneeDrive, &leg->footDrive};
  return t[i % 5];
}
static cassie_joint_out_t *joint_ptr(cassie_out_t *o, int i) {
  cassie_leg_out_t *leg = i < 3 ? &o->leftLeg : &o->rightLeg;
  cassie_joint_out_t *t[3] = {&leg->shinJoint, &leg->tarsusJoint, &leg->footJoint};
  return t[i % 3];
}
static void cassie_out_init_(cassie_out_t *o) { /* :695-734 and :666-692 */
  static const double tl[5] = {140.63, 140.63, 216.16, 216.16, 45.14}, gr[5] = {25, 25, 16, 16, 50};
  memset(o, 0, sizeof *o);
  o->isCalibrated = true;
  o->pelvis.medullaCounter = 1; o->pelvis.medullaCpuLoad = 159; o->pelvis.vtmTemperature = 40;
  o->pelvis.targetPc.etherCatStatus[1] = 8; o->pelvis.targetPc.etherCatStatus[4] = 1;
  o->pelvis.targetPc.taskExecutionTime = 2e-4; o->pelvis.targetPc.cpuTemperature = 60;
  o->pelvis.battery.dataGood = true; o->pelvis.battery.stateOfCharge = 1;
  for (int i = 0; i < 4; i++) o->pelvis.battery.temperature[i] = 30;
  for (int i = 0; i < 12; i++) o->pelvis.battery.voltage[i] = 4.2;
  o->pelvis.radio.radioReceiverSignalGood = true; o->pelvis.radio.receiverMedullaSignalGood = true; o->pelvis.radio.channel[8] = 1;
  o->pelvis.vectorNav.dataGood = true; o->pelvis.vectorNav.pressure = 101.325; o->pelvis.vectorNav.temperature = 25;
  for (int l = 0; l < 2; l++) {
    cassie_leg_out_t *leg = l ? &o->rightLeg : &o->leftLeg; leg->medullaCounter = 1; leg->medullaCpuLoad = 94;
    for (int i = 0; i < 5; i++) { elmo_out_t *e = drive_ptr(o, 5 * l + i); e->statusWord = 0x0637; e->dcLinkVoltage = 48; e->driveTemperature = 30; e->torqueLimit = tl[i]; e->gearRatio = gr[i]; }
  }
}

/* ---- Agility-block twins (closed source in the reference; semantics from SURVEY.md 8a-2 / 8a-3) */
static void o_task_pd_leg(int side, const pd_task_in_t *t, const double ang[7], const double rate[7], double tq[5]);
void o_pd_input_step(const pd_in_t *u, const cassie_out_t *o, double torque[10]) {
  for (int i = 0; i < 10; i++) {
    const pd_motor_in_t *p = i < 5 ? &u->leftLeg.motorPd : &u->rightLeg.motorPd; int k = i % 5;
    const elmo_out_t *e = drive_ptr((cassie_out_t *)o, i);
    torque[i] = p->torque[k] + p->pGain[k] * (p->pTarget[k] - e->position) + p->dGain[k] * (p->dTarget[k] - e->velocity);
  }
  for (int s = 0; s < 2; s++) {
    double ang[7], rate[7]; cassie_out_t *oo = (cassie_out_t *)o;
    for (int i = 0; i < 4; i++) { ang[i] = drive_ptr(oo, 5 * s + i)->position; rate[i] = drive_ptr(oo, 5 * s + i)->velocity; }
    ang[4] = joint_ptr(oo, 3 * s)->position; rate[4] = joint_ptr(oo, 3 * s)->velocity; ang[5] = joint_ptr(oo, 3 * s + 1)->position; rate[5] = joint_ptr(oo, 3 * s + 1)->velocity;
    ang[6] = drive_ptr(oo, 5 * s + 4)->position; rate[6] = drive_ptr(oo, 5 * s + 4)->velocity;
    o_task_pd_leg(s, s ? &u->rightLeg.taskPd : &u->leftLeg.taskPd, ang, rate, torque + 5 * s);
  }
}
/* soft joint limits = hard limits shrunk by W = 0.15 rad; order hipRoll hipYaw hipPitch knee foot, left leg then right.
 * Values recovered by bisection on the archive (tests/golden/make_golden.py documents the probe): hard limits in degrees are
 * hipRoll [-15, 20] (mirrored on the right), hipYaw +-22, hipPitch [-50, 80], knee [-156, -42], foot [-140, -35]. */
#define DEG (M_PI / 180.0)
static const double core_lo_deg[10] = {-15, -22, -50, -156, -140, -20, -22, -50, -156, -140};
static const double core_hi_deg[10] = {20, 22, 80, -42, -35, 15, 22, 80, -42, -35};
static const double core_K[5] = {1000, 800, 1200, 1200, 100}, core_C[5] = {12, 12, 36, 36, 7};
void o_core_sim_step(const double u[10], const cassie_out_t *o, double out[10]) {
  double pos[10], vel[10], lim[10], add[10] = {0}, scale = 1.0; const double W = 0.15;
  for (int i = 0; i < 10; i++) { const elmo_out_t *e = drive_ptr((cassie_out_t *)o, i); pos[i] = e->position; vel[i] = e->velocity; lim[i] = e->torqueLimit; }
  for (int i = 0; i < 10; i++) {
    int k = i % 5; double dhi = pos[i] - (core_hi_deg[i] * DEG - W), dlo = (core_lo_deg[i] * DEG + W) - pos[i];
    if (dhi > 0) { add[i] -= core_K[k] * dhi * (1 + dhi / W) + core_C[k] * fmin(dhi / W, 1.0) * vel[i]; scale *= fmax(0.0, 1 - dhi / W); }
    if (dlo > 0) { add[i] += core_K[k] * dlo * (1 + dlo / W) - core_C[k] * fmin(dlo / W, 1.0) * vel[i]; scale *= fmax(0.0, 1 - dlo / W); }
  }
  for (int l = 0; l < 2; l++) { /* coupled row: hipPitch + knee >= -135 deg; each joint damped by its own velocity */
    int a = 5 * l + 2, b = 5 * l + 3; double dsum = -135 * DEG - (pos[a] + pos[b]);
    if (dsum > 0) {
      double t = 1200 * dsum * (1 + dsum / W);
      double r = fmin(dsum / W, 1.0); add[a] += t - 36 * r * vel[a]; add[b] += t - 36 * r * vel[b]; scale *= fmax(0.0, 1 - dsum / W);
    }
  }
  int sto = !(o->pelvis.radio.channel[8] >= 1);
  for (int i = 0; i < 10; i++) { double t = sto ? 0.0 : u[i] * scale + add[i]; out[i] = clampd(t, -lim[i], lim[i]); }
}

/* ---- state_output_step: the STATELESS part of the closed estimator, decoded by probing the archive (oracle/probe_estimator.c,
 * oracle/probe_est_tools.py; agreement with the archive 1e-14 over random inputs, tests/test_agility_twins.py):
 *  - pass-through: motor / joint position, velocity, torque, radio, battery, IMU gyro;
 *  - pelvis.orientation = the IMU quaternion sent through its rotation matrix and back (mat2quat with MuJoCo's branches): +q or -q, the sign
 *    being the one that makes w positive when the trace is positive, else the largest-diagonal component positive;
 *  - pelvis.translationalAcceleration = accelerometer - R(q)' (0, 0, 9.806) - w x (w x r)   (body frame, gravity removed, moved from the
 *    IMU site to the pelvis origin without the angular-acceleration term);
 *  - per foot: forward kinematics of a 7-link chain pelvis -> hipRoll -> hipYaw -> hipPitch -> knee -> shin -> tarsus -> foot with the MJCF's
 *    offsets (model/cassie.xml:96-152) and EXACT quarter turns, driven by the MEASURED angles (motor encoders for the five drives, joint
 *    encoders for shin and tarsus; the foot joint encoder is not used).  position = the point (0.01762, 0.05219, 0) of the foot frame in
 *    the pelvis frame; orientation = foot frame turned by a fixed rotation (40 degrees), as a quaternion (mat2quat branches as MuJoCo's);
 *    footRotationalVelocity / footTranslationalVelocity = Jacobian times the measured rates, expressed in THAT FOOT FRAME.
 *  - toeForce = heelForce: o_est_leg_force below (spring torques through the closed four-bar; single-precision agreement only).
 * The STATEFUL part (o_est_filter_step below) was decoded from the block's own memory (oracle/probe_estimator.c probe_est_mem: the block
 * keeps its states, covariances and noise matrices in plain doubles) and agrees with the archive to 1e-15 when fed the archive's own
 * stateless outputs (tests/test_estimator_filter.py): pelvis.position / translationalVelocity / externalForce and terrain.height.
 * externalMoment and terrain.slope are always zero in the archive too. */
static void est_mat2quat(double *q, const double *R) { /* R row-major */
  double t = R[0] + R[4] + R[8];
  if (t > 0) { double s = sqrt(t + 1) * 2; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { double s = sqrt(1 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { double s = sqrt(1 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s; }
  else { double s = sqrt(1 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s; }
}
static void o_leg_chain(int side, const double ang[7], double pos[3], double Rf[9], double axis[7][3], double anchor[7][3]) {
  /* ang: hipRoll, hipYaw, hipPitch, knee, shin, tarsus, foot; pelvis frame */
  const double sg = side ? -1.0 : 1.0;
  const double off[7][3] = {{0.021, 0.135 * sg, 0}, {0, 0, -0.07}, {0, 0, -0.09}, {0.12, 0, 0.0045 * sg}, {0.06068, 0.04741, 0}, {0.43476, 0.02, 0}, {0.408, -0.04, 0}};
  /* fixed frame of each link in its parent (columns = x, y, z axes): quarter turns for the three hip links, identity below */
  static const double F[3][9] = {{0, 0, 1, 0, 1, 0, -1, 0, 0}, {0, 0, -1, 0, 1, 0, 1, 0, 0}, {0, 1, 0, 0, 0, -1, -1, 0, 0}};
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  for (int i = 0; i < 7; i++) {
    double v[3], T[9], c = cos(ang[i]), s = sin(ang[i]);
    mulMatVec3(v, R, off[i]); for (int k = 0; k < 3; k++) p[k] += v[k];
    if (i < 3) { mulMatMat3(T, R, F[i]); copyv(R, T, 9); }
    const double Z[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
    mulMatMat3(T, R, Z); copyv(R, T, 9);
    copyv(anchor[i], p, 3); axis[i][0] = R[2]; axis[i][1] = R[5]; axis[i][2] = R[8];
  }
  const double c40 = cos(40 * M_PI / 180), s40 = sin(40 * M_PI / 180), Roff[9] = {-c40, 0, -s40, s40, 0, -c40, 0, -1, 0}, loc[3] = {0.01762, 0.05219, 0};
  double v[3];
  mulMatMat3(Rf, R, Roff); mulMatVec3(v, R, loc); for (int k = 0; k < 3; k++) pos[k] = p[k] + v[k];
}
void o_est_foot(int side, const double ang[7], const double rate[7], double pos[3], double quat[4], double rotvel[3], double linvel[3]) {
  double Rf[9], axis[7][3], anchor[7][3], w[3] = {0, 0, 0}, vl[3] = {0, 0, 0};
  o_leg_chain(side, ang, pos, Rf, axis, anchor);
  est_mat2quat(quat, Rf);
  for (int i = 0; i < 7; i++) {
    double d[3] = {pos[0] - anchor[i][0], pos[1] - anchor[i][1], pos[2] - anchor[i][2]}, cr[3];
    cross(cr, axis[i], d);
    for (int k = 0; k < 3; k++) { w[k] += axis[i][k] * rate[i]; vl[k] += cr[k] * rate[i]; }
  }
  mulMatTVec3(rotvel, Rf, w); mulMatTVec3(linvel, Rf, vl);
}
/* pd_input_step, taskPd branch (closed; decoded by probing the archive, oracle/probe_estimator.c probe_pd; pinned in tests/test_agility_twins.py).
 * Per leg, six task coordinates x = [foot position (pelvis frame, the estimator's point), yaw, pitch, roll of the estimator's foot frame (ZYX Euler
 * angles of its quaternion)], their rates v = [footTranslationalVelocity, footRotationalVelocity] exactly as the estimator reports them (foot
 * frame), w_k = torque_k + pGain_k (pTarget_k - x_k) + dGain_k (dTarget_k - v_k), and motor torques += A' w with A = [Jw; Jv], the angular then the
 * linear Jacobian of the foot point (pelvis frame) with respect to the five MOTOR angles, shin / tarsus held at their measured values.  (Component
 * k of w multiplies row k of A -- position errors meet the angular rows; that is what the archive computes.) */
static void o_task_pd_leg(int side, const pd_task_in_t *t, const double ang[7], const double rate[7], double tq[5]) {
  double pos[3], Rf[9], axis[7][3], anchor[7][3], quat[4], rot[3], lin[3], x[6], v[6];
  static const int mot[5] = {0, 1, 2, 3, 6};
  int any = 0; for (int k = 0; k < 6; k++) any |= (t->torque[k] != 0 || t->pGain[k] != 0 || t->dGain[k] != 0);
  if (!any) return;
  o_leg_chain(side, ang, pos, Rf, axis, anchor);
  o_est_foot(side, ang, rate, pos, quat, rot, lin);
  const double qw = quat[0], qx = quat[1], qy = quat[2], qz = quat[3];
  x[0] = pos[0]; x[1] = pos[1]; x[2] = pos[2];
  x[3] = atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz)); x[4] = asin(2 * (qw * qy - qz * qx)); x[5] = atan2(2 * (qw * qx + qy * qz), 1 - 2 * (qx * qx + qy * qy));
  for (int k = 0; k < 3; k++) { v[k] = lin[k]; v[3 + k] = rot[k]; }
  double w[6]; for (int k = 0; k < 6; k++) w[k] = t->torque[k] + t->pGain[k] * (t->pTarget[k] - x[k]) + t->dGain[k] * (t->dTarget[k] - v[k]);
  for (int j = 0; j < 5; j++) {
    const int i = mot[j]; double d[3] = {pos[0] - anchor[i][0], pos[1] - anchor[i][1], pos[2] - anchor[i][2]}, cr[3];
    cross(cr, axis[i], d);
    tq[j] += axis[i][0] * w[0] + axis[i][1] * w[1] + axis[i][2] * w[2] + cr[0] * w[3] + cr[1] * w[4] + cr[2] * w[5];
  }
}
/* toe / heel force of one leg (state_output_step, closed; decoded by fitting the archive, oracle/probe_fit_forces.py): the two leaf springs'
 * torques mapped to a force at the foot point through the closed four-bar.  Both outputs are the same vector:
 *   toeForce = heelForce = Rz(yaw)' R(q) [f_x, 0, f_z],   (f_x, f_z) = -1/2 J_c^-T [1500 shin; 1250 (H - 2.586e-6)]
 * H = heel-spring angle that closes the achilles rod: |A - B(knee, shin, tarsus, H)| = 0.5012 with A = (0, 0, 0.045) on the hip-pitch link, the
 * heel spring mounted on the tarsus as in the MJCF (model/cassie.xml:132) and the rod end at (0.11877, -0.01, 0) of the heel-spring frame;
 * J_c = [p_s - p_t b/a, p_t / a] (pelvis x and z rows), p_s / p_t the serial-chain partials of the foot point w.r.t. shin / tarsus and
 * a = dH/dtarsus, b = dH/dshin from the closure; q the IMU quaternion, yaw its ZYX heading.  The archive evaluates this in single precision:
 * agreement is 4e-3 N on forces of 150 N, not the 1e-12 of the kinematic outputs. */
void o_est_leg_force(int side, const double ang[7], const double quat[4], double force[3]) {
  const double sg = side ? -1.0 : 1.0, kn = ang[3], sh = ang[4], ta = ang[5];
  /* planar chain in the hip-pitch frame (z = the common joint axis) */
  const double A[3] = {0, 0, 0.045 * sg}, k0[3] = {0.12, 0, 0.0045 * sg}, hsp[3] = {-0.01269, -0.03059, 0.00092 * sg}, Bl[3] = {0.11877, -0.01, 0};
  double hx[3] = {-0.91211, 0.40829, 0.036948 * sg}, hy[3] = {-0.40992, -0.90952, -0.068841 * sg}, hz[3];
  { double n = sqrt(dot3(hx, hx)); for (int k = 0; k < 3; k++) hx[k] /= n; double d = dot3(hx, hy); for (int k = 0; k < 3; k++) hy[k] -= d * hx[k];
    n = sqrt(dot3(hy, hy)); for (int k = 0; k < 3; k++) hy[k] /= n; cross(hz, hx, hy); }      /* xyaxes -> frame, as the MJCF compiler does */
  const double c1 = cos(kn), s1 = sin(kn), c2 = cos(kn + sh), s2 = sin(kn + sh), c3 = cos(kn + sh + ta), s3 = sin(kn + sh + ta);
  const double s0[3] = {k0[0] + c1 * 0.06068 - s1 * 0.04741, k0[1] + s1 * 0.06068 + c1 * 0.04741, k0[2]};
  const double t0[3] = {s0[0] + c2 * 0.43476 - s2 * 0.02, s0[1] + s2 * 0.43476 + c2 * 0.02, s0[2]};
  const double R3[9] = {c3, -s3, 0, s3, c3, 0, 0, 0, 1};
  double hs0[3], v[3], HF[9] = {hx[0], hy[0], hz[0], hx[1], hy[1], hz[1], hx[2], hy[2], hz[2]}, RH[9], axh[3];
  mulMatVec3(v, R3, hsp); for (int k = 0; k < 3; k++) hs0[k] = t0[k] + v[k];
  mulMatMat3(RH, R3, HF); axh[0] = RH[2]; axh[1] = RH[5]; axh[2] = RH[8];                      /* heel-spring joint axis */
  double H = 0, B[3], dB[3], gd = 1;
  for (int it = 0; it < 8; it++) {                                                              /* Newton on g(H) = |B - A|^2 - L^2 */
    const double ch = cos(H), shh = sin(H), Zl[3] = {ch * Bl[0] - shh * Bl[1], shh * Bl[0] + ch * Bl[1], Bl[2]};
    mulMatVec3(v, RH, Zl); for (int k = 0; k < 3; k++) { B[k] = hs0[k] + v[k]; dB[k] = B[k] - A[k]; }
    double r[3] = {B[0] - hs0[0], B[1] - hs0[1], B[2] - hs0[2]}, dBd[3]; cross(dBd, axh, r);
    const double g = dot3(dB, dB) - 0.5012 * 0.5012; gd = 2 * dot3(dB, dBd);
    H -= g / gd;
  }
  { const double ch = cos(H), shh = sin(H), Zl[3] = {ch * Bl[0] - shh * Bl[1], shh * Bl[0] + ch * Bl[1], Bl[2]};
    mulMatVec3(v, RH, Zl); for (int k = 0; k < 3; k++) { B[k] = hs0[k] + v[k]; dB[k] = B[k] - A[k]; }
    double r[3] = {B[0] - hs0[0], B[1] - hs0[1], B[2] - hs0[2]}, dBd[3]; cross(dBd, axh, r); gd = 2 * dot3(dB, dBd); }
  const double ez[3] = {0, 0, 1}; double rs[3] = {B[0] - s0[0], B[1] - s0[1], B[2] - s0[2]}, rt[3] = {B[0] - t0[0], B[1] - t0[1], B[2] - t0[2]}, dBs[3], dBt[3];
  cross(dBs, ez, rs); cross(dBt, ez, rt);
  const double a = -2 * dot3(dB, dBt) / gd, b = -2 * dot3(dB, dBs) / gd;
  /* foot-point partials in the pelvis frame from the serial chain */
  double pos[3], Rf[9], axis[7][3], anchor[7][3], d4[3], d5[3], ps[3], pt[3];
  o_leg_chain(side, ang, pos, Rf, axis, anchor);
  for (int k = 0; k < 3; k++) { d4[k] = pos[k] - anchor[4][k]; d5[k] = pos[k] - anchor[5][k]; }
  cross(ps, axis[4], d4); cross(pt, axis[5], d5);
  /* J_c' f = tau with J_c = [ps - pt b/a, pt/a] (x and z rows) */
  const double j00 = ps[0] - pt[0] * b / a, j10 = ps[2] - pt[2] * b / a, j01 = pt[0] / a, j11 = pt[2] / a;
  const double t0_ = 1500.0 * sh, t1_ = 1250.0 * (H - 2.586e-6), det = j00 * j11 - j10 * j01;
  const double fx = -0.5 * (t0_ * j11 - t1_ * j10) / det, fz = -0.5 * (-t0_ * j01 + t1_ * j00) / det;
  /* heading-free world frame */
  double R[9]; quat2Mat(R, quat);
  const double yaw = atan2(2 * (quat[0] * quat[3] + quat[1] * quat[2]), 1 - 2 * (quat[2] * quat[2] + quat[3] * quat[3])), cy = cos(yaw), sy = sin(yaw);
  const double wv[3] = {R[0] * fx + R[2] * fz, R[3] * fx + R[5] * fz, R[6] * fx + R[8] * fz};
  force[0] = cy * wv[0] + sy * wv[1]; force[1] = -sy * wv[0] + cy * wv[1]; force[2] = wv[2];
}

/* ---- the estimator's stateful part: three per-axis Kalman filters in the world frame, 2 kHz (dt = 0.0005), nominal mass 31 kg, g = 9.806.
 * Inputs per call (all outputs of the stateless part): R = R(orientation), the foot points in the pelvis frame, the leg forces
 * (toeForce + heelForce, i.e. the leg's force on the ground: negative z when it carries load), translationalAcceleration.
 *   y_i = -R foot_i                      (pelvis relative to foot i, world frame), a = R translationalAcceleration
 *   f_i = min(F_i.z, 0);  contact = -(f_L + f_R) > 1 N;  w_meas = contact ? f_L / (f_L + f_R) : 0.5;  q_i = F_i.z < -50 N ? 1e-10 : 1e-6
 * x and y axes: extended filter, state [p, v, footL, footR, w, F_ext], P0 = 1e-6 I, Q = diag(1e-8, 1e-8, q_L, q_R, 1e-5, 1e-2),
 *   predict  p += dt v;  in contact (linear inverted pendulum of height 1 m over the load-weighted foot point):
 *            v += dt g (p - w footL - (1 - w) footR) + dt/m F_ext      (Jacobian row [c, 1, -c w, -c (1-w), -c (footL - footR), dt/m], c = dt g);
 *            out of contact v is held (row [0, 1, 0, 0, 0, 0]);
 *   measure  [p - footL, p - footR, w, v] = [y_L, y_R, w_meas, v_previous + dt a],  R = diag(1e-6, 1e-6, 1e-6, 1).
 * z axis: linear filter, state [p, v, footL, footR, F_ext], Q = diag(1e-8, 1e-8, q_L, q_R, 1e-2),
 *   predict  p += dt v;  v += dt (-g - (f_L + f_R) / m) + dt/m F_ext;   measure [p - footL, p - footR] = [y_L.z, y_R.z], R = 1e-6 I.
 * First call: states start at [0, 0, y_L, y_R, 0.5, 0] (x, y) and [0, 0, y_L.z, y_R.z, m g] (z) - the feet at +y, as the archive has it.
 * terrain.height: first-order lag (1 s, backward Euler) of p_z - (w_meas y_L.z + (1 - w_meas) y_R.z), advanced only in contact. */
void o_est_filter_reset(OEst *e) { memset(e, 0, sizeof *e); }
/* x <- x + K (zm - H x), P <- (I - K H) P for n states, k measurements (H row-major k x n, R diagonal) */
static void kf_update_(int n, int k, double *x, double *P, const double *H, const double *Rd, const double *zm) {
  double PHt[6 * 4], S[4 * 4], Si[4 * 4], K[6 * 4], inn[4], Pn[36];
  for (int i = 0; i < n; i++) for (int j = 0; j < k; j++) { double a = 0; for (int l = 0; l < n; l++) a += P[i * n + l] * H[j * n + l]; PHt[i * k + j] = a; }
  for (int i = 0; i < k; i++) for (int j = 0; j < k; j++) { double a = i == j ? Rd[i] : 0; for (int l = 0; l < n; l++) a += H[i * n + l] * PHt[l * k + j]; S[i * k + j] = a; }
  /* inverse by Gauss-Jordan with partial pivoting */
  double M[4][8];
  for (int i = 0; i < k; i++) for (int j = 0; j < k; j++) { M[i][j] = S[i * k + j]; M[i][k + j] = i == j; }
  for (int c = 0; c < k; c++) {
    int pv = c; for (int r = c + 1; r < k; r++) if (fabs(M[r][c]) > fabs(M[pv][c])) pv = r;
    if (pv != c) for (int j = 0; j < 2 * k; j++) { double t = M[c][j]; M[c][j] = M[pv][j]; M[pv][j] = t; }
    double d = M[c][c]; for (int j = 0; j < 2 * k; j++) M[c][j] /= d;
    for (int r = 0; r < k; r++) if (r != c) { double f = M[r][c]; for (int j = 0; j < 2 * k; j++) M[r][j] -= f * M[c][j]; }
  }
  for (int i = 0; i < k; i++) for (int j = 0; j < k; j++) Si[i * k + j] = M[i][k + j];
  for (int i = 0; i < n; i++) for (int j = 0; j < k; j++) { double a = 0; for (int l = 0; l < k; l++) a += PHt[i * k + l] * Si[l * k + j]; K[i * k + j] = a; }
  for (int j = 0; j < k; j++) { double a = zm[j]; for (int l = 0; l < n; l++) a -= H[j * n + l] * x[l]; inn[j] = a; }
  for (int i = 0; i < n; i++) for (int j = 0; j < k; j++) x[i] += K[i * k + j] * inn[j];
  double HP[4 * 6];   /* (I - K H) P with H P formed as such (not as (P H')'): the archive's form; P is only symmetric to rounding */
  for (int l = 0; l < k; l++) for (int j = 0; j < n; j++) { double a = 0; for (int q = 0; q < n; q++) a += H[l * n + q] * P[q * n + j]; HP[l * n + j] = a; }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double a = P[i * n + j]; for (int l = 0; l < k; l++) a -= K[i * k + l] * HP[l * n + j]; Pn[i * n + j] = a; }
  memcpy(P, Pn, sizeof(double) * n * n);
}
/* P <- A P A' + diag(Qd) */
static void kf_predict_cov_(int n, double *P, const double *A, const double *Qd) {
  double T[36], Pn[36];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double a = 0; for (int l = 0; l < n; l++) a += A[i * n + l] * P[l * n + j]; T[i * n + j] = a; }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double a = i == j ? Qd[i] : 0; for (int l = 0; l < n; l++) a += T[i * n + l] * A[j * n + l]; Pn[i * n + j] = a; }
  memcpy(P, Pn, sizeof(double) * n * n);
}
void o_est_filter_step(OEst *e, state_out_t *y) {
  const double dt = 0.0005, mass = 31, grav = 9.806, c = dt * grav;
  double R[9], yL[3], yR[3], aw[3];
  quat2Mat(R, y->pelvis.orientation);
  mulMatVec3(yL, R, y->leftFoot.position); mulMatVec3(yR, R, y->rightFoot.position); mulMatVec3(aw, R, y->pelvis.translationalAcceleration);
  for (int k = 0; k < 3; k++) { yL[k] = -yL[k]; yR[k] = -yR[k]; }
  const double FLz = y->leftFoot.toeForce[2] + y->leftFoot.heelForce[2], FRz = y->rightFoot.toeForce[2] + y->rightFoot.heelForce[2];
  const double fl = FLz < 0 ? FLz : 0, fr = FRz < 0 ? FRz : 0;
  const int contact = -(fl + fr) > 1.0;
  const double wm = contact ? fl / (fl + fr) : 0.5, qL = FLz < -50 ? 1e-10 : 1e-6, qR = FRz < -50 ? 1e-10 : 1e-6;
  if (!e->started) {
    for (int ax = 0; ax < 2; ax++) {
      const double x0[6] = {0, 0, yL[ax], yR[ax], 0.5, 0}; copyv(e->x[ax], x0, 6);
      memset(e->P[ax], 0, sizeof e->P[ax]); for (int i = 0; i < 6; i++) e->P[ax][7 * i] = 1e-6;
    }
    const double z0[5] = {0, 0, yL[2], yR[2], mass * grav}; copyv(e->z, z0, 5);
    memset(e->Pz, 0, sizeof e->Pz); for (int i = 0; i < 5; i++) e->Pz[6 * i] = 1e-6;
    e->terrain = 0; e->started = 1;
  }
  static const double H6[24] = {1, 0, -1, 0, 0, 0,  1, 0, 0, -1, 0, 0,  0, 0, 0, 0, 1, 0,  0, 1, 0, 0, 0, 0}, R4[4] = {1e-6, 1e-6, 1e-6, 1};
  for (int ax = 0; ax < 2; ax++) {
    double *x = e->x[ax], A[36] = {0}, zm[4] = {yL[ax], yR[ax], wm, x[1] + dt * aw[ax]};
    const double Qd[6] = {1e-8, 1e-8, qL, qR, 1e-5, 1e-2}, p0 = x[0], v0 = x[1], w = x[4];
    for (int i = 0; i < 6; i++) A[7 * i] = 1;
    A[1] = dt;
    if (contact) { A[6] = c; A[8] = -c * w; A[9] = -c * (1 - w); A[10] = -c * (x[2] - x[3]); A[11] = dt / mass; x[1] = v0 + c * (p0 - w * x[2] - (1 - w) * x[3]) + dt / mass * x[5]; }
    x[0] = p0 + dt * v0;
    kf_predict_cov_(6, e->P[ax], A, Qd);
    kf_update_(6, 4, x, e->P[ax], H6, R4, zm);
  }
  {
    static const double H5[10] = {1, 0, -1, 0, 0,  1, 0, 0, -1, 0}, R2[2] = {1e-6, 1e-6};
    double *x = e->z, A[25] = {0}, zm[2] = {yL[2], yR[2]};
    const double Qd[5] = {1e-8, 1e-8, qL, qR, 1e-2}, p0 = x[0], v0 = x[1];
    for (int i = 0; i < 5; i++) A[6 * i] = 1;
    A[1] = dt; A[9] = dt / mass;
    x[0] = p0 + dt * v0; x[1] = v0 + dt / mass * x[4] + dt * (-grav - (fl + fr) / mass);
    kf_predict_cov_(5, e->Pz, A, Qd);
    kf_update_(5, 2, x, e->Pz, H5, R2, zm);
  }
  if (contact) e->terrain = (e->terrain + dt * (e->z[0] - (wm * yL[2] + (1 - wm) * yR[2]))) / (1 + dt);
  for (int ax = 0; ax < 2; ax++) { y->pelvis.position[ax] = e->x[ax][0]; y->pelvis.translationalVelocity[ax] = e->x[ax][1]; y->pelvis.externalForce[ax] = e->x[ax][5]; }
  y->pelvis.position[2] = e->z[0]; y->pelvis.translationalVelocity[2] = e->z[1]; y->pelvis.externalForce[2] = e->z[4];
  y->terrain.height = e->terrain;
}
void o_state_output_step(const cassie_out_t *o, state_out_t *y) {
  cassie_out_t out = *o;
  memset(y, 0, sizeof *y);
  for (int i = 0; i < 10; i++) { elmo_out_t *e = drive_ptr(&out, i); y->motor.position[i] = e->position; y->motor.velocity[i] = e->velocity; y->motor.torque[i] = e->torque; }
  for (int i = 0; i < 6; i++) { cassie_joint_out_t *j = joint_ptr(&out, i); y->joint.position[i] = j->position; y->joint.velocity[i] = j->velocity; }
  const double *q = out.pelvis.vectorNav.orientation;
  { double Rq[9]; quat2Mat(Rq, q); est_mat2quat(y->pelvis.orientation, Rq); }   /* through the rotation matrix: +-q with mat2quat's sign choice */
  copyv(y->pelvis.rotationalVelocity, out.pelvis.vectorNav.angularVelocity, 3);
  { /* accelerometer moved from the IMU (r = (0.03155, 0, -0.079996) from the pelvis origin; the MJCF site has -0.07996, model/cassie.xml:87) to the pelvis origin, centripetal part only: a - w x (w x r) */
    double R[9], wr[3], wwr[3]; const double r[3] = {0.03155, 0, -0.079996}, *w = out.pelvis.vectorNav.angularVelocity;
    quat2Mat(R, q); cross(wr, w, r); cross(wwr, w, wr);
    for (int k = 0; k < 3; k++) y->pelvis.translationalAcceleration[k] = out.pelvis.vectorNav.linearAcceleration[k] - R[6 + k] * 9.806 - wwr[k]; }
  for (int s = 0; s < 2; s++) {
    state_foot_out_t *f = s ? &y->rightFoot : &y->leftFoot; double ang[7], rate[7];
    for (int i = 0; i < 4; i++) { ang[i] = y->motor.position[5 * s + i]; rate[i] = y->motor.velocity[5 * s + i]; }
    ang[4] = y->joint.position[3 * s]; rate[4] = y->joint.velocity[3 * s]; ang[5] = y->joint.position[3 * s + 1]; rate[5] = y->joint.velocity[3 * s + 1];
    ang[6] = y->motor.position[5 * s + 4]; rate[6] = y->motor.velocity[5 * s + 4];
    o_est_foot(s, ang, rate, f->position, f->orientation, f->footRotationalVelocity, f->footTranslationalVelocity);
    o_est_leg_force(s, ang, q, f->toeForce); copyv(f->heelForce, f->toeForce, 3);
  }
  copyv(y->radio.channel, out.pelvis.radio.channel, 16); y->radio.signalGood = true; y->battery.stateOfCharge = out.pelvis.battery.stateOfCharge;
}

/* stateless part then the filters (one OEst per estimator instance) */
void o_state_output_step_full(OEst *e, const cassie_out_t *o, state_out_t *y) { o_state_output_step(o, y); o_est_filter_step(e, y); }
OEst *o_est_new(void) { return calloc(1, sizeof(OEst)); }
void o_est_free(OEst *e) { free(e); }

OSim *osim_new(const char *model_path) {
  OSim *c = calloc(1, sizeof(OSim));
  c->m = omodel_load(model_path); if (!c->m) { free(c); return NULL; }
  c->d = calloc(1, sizeof(OData));
  cassie_out_init_(&c->cassie_out);
  copyv(c->d->qpos, c->m->qpos0, c->m->nq);
  static const double qi[28] = {0.0045, 0, 0.4973, 0.9785, -0.0164, 0.01787, -0.2049, -1.1997, 0, 1.4267, 0, -1.5244, 1.5244, -1.5968,
                                -0.0045, 0, 0.4973, 0.9786, 0.00386, -0.01524, -0.2051, -1.1997, 0, 1.4267, 0, -1.5244, 1.5244, -1.5968}; /* :1023-1027 */
  copyv(c->d->qpos + 7, qi, 28);
  o_forward(c->m, c->d); /* mj_forward at :1029 */
#ifdef ORACLE_USE_AGILITY_REF
  c->core = cassie_core_sim_alloc(); c->est = state_output_alloc(); c->pd = pd_input_alloc();
  cassie_core_sim_setup(c->core); state_output_setup(c->est); pd_input_setup(c->pd);
#endif
  return c;
}
void osim_free(OSim *c) {
  if (!c) return;
#ifdef ORACLE_USE_AGILITY_REF
  cassie_core_sim_free(c->core); state_output_free(c->est); pd_input_free(c->pd);
#endif
  free(c->d->efc_AR); free(c->d->efc_JM2); free(c->d); omodel_free(c->m); free(c);
}

static double motor_(OSim *c, int i, double u, int sto) { /* :638-664 */
  const OModel *m = c->m; OData *d = c->d;
  double ratio = m->actuator_gear[i], tmax = m->actuator_ctrlrange[i][1], w = d->actuator_velocity[i], wmax = m->actuator_user[i] * 2 * M_PI / 60;
  double tlim = 2 * tmax * (1 - fabs(w) / wmax); tlim = fmax(fmin(tlim, tmax), 0);
  if (sto) u = 0;
  double tau = copysign(fmin(fabs(u / ratio), tlim), u);
  d->ctrl[i] = c->torque_delay[i][DELAY - 1];
  for (int k = DELAY - 1; k > 0; k--) c->torque_delay[i][k] = c->torque_delay[i][k - 1];
  c->torque_delay[i][0] = tau;
  return d->ctrl[i] * ratio;
}
static void sensor_data_(OSim *c) { /* :737-774, 558-635 */
  static const int dsid[10] = {0, 1, 2, 3, 4, 8, 9, 10, 11, 12}, jsid[6] = {5, 6, 7, 13, 14, 15};
  const OModel *m = c->m; const double *sd = c->d->sensordata;
  for (int i = 0; i < NUM_DRIVES; i++) {
    elmo_out_t *dr = drive_ptr(&c->cassie_out, i); int s = dsid[i], bits = (int)m->sensor_user[s], *x = c->drive_filter[i];
    int enc = (int)(sd[s] / (2 * M_PI) * (1 << bits));
    double ratio = m->actuator_gear[m->sensor_objid[s]], scale = (2 * M_PI) / (1 << bits) / ratio;
    dr->position = enc * scale;
    int allzero = 1; for (int k = 0; k < 9; k++) allzero &= x[k] == 0;
    if (allzero) for (int k = 0; k < 9; k++) x[k] = enc;
    for (int k = 8; k > 0; k--) x[k] = x[k - 1];
    x[0] = enc;
    int y = 0; for (int k = 0; k < 9; k++) y += x[k] * drive_b[k];
    dr->velocity = y * scale / M_PI;
  }
  for (int i = 0; i < NUM_JOINTS; i++) {
    cassie_joint_out_t *jn = joint_ptr(&c->cassie_out, i); int s = jsid[i], bits = (int)m->sensor_user[s]; double *x = c->joint_filter_x[i], *y = c->joint_filter_y[i];
    int enc = (int)(sd[s] / (2 * M_PI) * (1 << bits)); double scale = (2 * M_PI) / (1 << bits);
    jn->position = enc * scale;
    int allzero = 1; for (int k = 0; k < 4; k++) allzero &= x[k] == 0;
    if (allzero) for (int k = 0; k < 4; k++) x[k] = jn->position;
    for (int k = 3; k > 0; k--) x[k] = x[k - 1];
    x[0] = jn->position;
    for (int k = 2; k > 0; k--) y[k] = y[k - 1];
    y[0] = 0;
    for (int k = 0; k < 4; k++) y[0] += x[k] * joint_b[k];
    for (int k = 1; k < 3; k++) y[0] -= y[k] * joint_a[k];
    jn->velocity = y[0];
  }
  vectornav_out_t *vn = &c->cassie_out.pelvis.vectorNav;
  copyv(vn->orientation, sd + 16, 4); copyv(vn->angularVelocity, sd + 20, 3); copyv(vn->linearAcceleration, sd + 23, 3); copyv(vn->magneticField, sd + 26, 3);
}

/* cassie_sim_step_pd, src/cassiemujoco.c:1147-1157 (with :1137-1145 and :1115-1135 inlined).
 * y may be NULL.  cassie_out_copy (optional) receives the cassie_out_t the reference hands to the estimator. */
static int g_no2khz = 0;   /* osim_step_pd_no2khz: one mj_step per call (src/cassiemujoco.c:1175) instead of round(5e-4 / timestep) */
void osim_step_pd(OSim *c, const pd_in_t *u, state_out_t *y, cassie_out_t *cassie_out_copy) {
  double tq_user[10], tq_in[10];
#ifdef ORACLE_USE_AGILITY_REF
  cassie_user_in_t ui; cassie_in_t ci;
  pd_input_step(c->pd, u, &c->cassie_out, &ui);
  cassie_core_sim_step(c->core, &ui, &c->cassie_out, &ci);
  for (int i = 0; i < 10; i++) { const cassie_leg_in_t *leg = i < 5 ? &ci.leftLeg : &ci.rightLeg; const elmo_in_t *t[5] = {&leg->hipRollDrive, &leg->hipYawDrive, &leg->hipPitchDrive, &leg->kneeDrive, &leg->footDrive}; tq_in[i] = t[i % 5]->torque; }
  (void)tq_user;
#else
  o_pd_input_step(u, &c->cassie_out, tq_user);
  o_core_sim_step(tq_user, &c->cassie_out, tq_in);
#endif
  int sto = c->cassie_out.pelvis.radio.channel[8] < 1;
  for (int i = 0; i < NUM_DRIVES; i++) drive_ptr(&c->cassie_out, i)->torque = motor_(c, i, tq_in[i], sto);
  sensor_data_(c);
  cassie_out_t out = c->cassie_out;
  if (cassie_out_copy) *cassie_out_copy = out;
  int mjsteps = g_no2khz ? 1 : (int)round(5e-4 / c->m->timestep);
  for (int i = 0; i < mjsteps; i++) o_step(c->m, c->d);
  if (y) {
#ifdef ORACLE_USE_AGILITY_REF
    state_output_step(c->est, &out, y);
#else
    o_state_output_step_full(&c->est2, &out, y); /* the decoded estimator */
#endif
  }
}

/* cassie_sim_step_pd_no2khz, src/cassiemujoco.c:1159-1181: the same blocks in the same order, ONE physics step */
void osim_step_pd_no2khz(OSim *c, const pd_in_t *u, state_out_t *y) { g_no2khz = 1; osim_step_pd(c, u, y, NULL); g_no2khz = 0; }

/* ------------------------------------------------------------------ derived-quantity queries (src/cassiemujoco.c:1586-1961)
 * Each function restates the reference function of the same name, INCLUDING which arrays it recomputes and which it reads stale:
 * after cassie_sim_step_pd the contact list, efc_force, xpos and cdof belong to the state the last sub-step started from, qvel / qpos are
 * one step newer. */
static void o_fwdPosition(const OModel *m, OData *d) { /* mj_fwdPosition [M] (camlight / tendon: nothing to do for these models) */
  o_kinematics(m, d); o_comPos(m, d);
  for (int i = 0; i < m->nu; i++) { int j = m->actuator_jntid[i]; d->actuator_length[i] = m->actuator_gear[i] * d->qpos[m->jnt_qposadr[j]]; }
  o_crb(m, d); copyv(d->qLD, d->qM, m->nM); factorI(m, d->qLD, d->qLDiagInv, d->qLDiagSqrtInv);
  o_collision(m, d); o_makeConstraint(m, d); o_projectConstraint(m, d);
}
static void o_subtreeVel(const OModel *m, OData *d) { /* mj_subtreeVel [M]; reads d->cvel as it stands */
  double body_vel[MAXB][6], dx[3], dv[3], dp[3], dL[3];
  for (int i = 0; i < m->nbody; i++) {
    /* mj_objectVelocity(mjOBJ_BODY, world orientation): cvel moved from the c-frame origin to the body's inertial frame origin */
    const double *cv = d->cvel[i]; double dif[3], cr[3];
    for (int k = 0; k < 3; k++) dif[k] = d->xipos[i][k] - d->subtree_com[m->body_rootid[i]][k];
    cross(cr, dif, cv);
    for (int k = 0; k < 3; k++) { body_vel[i][k] = cv[k]; body_vel[i][3 + k] = cv[3 + k] - cr[k]; }
    for (int k = 0; k < 3; k++) d->subtree_linvel[i][k] = body_vel[i][3 + k] * m->body_mass[i];
    mulMatTVec3(dv, d->ximat[i], body_vel[i]);
    for (int k = 0; k < 3; k++) dv[k] *= m->body_inertia[i][k];
    mulMatVec3(d->subtree_angmom[i], d->ximat[i], dv);
  }
  for (int i = m->nbody - 1; i >= 0; i--) {
    if (i) for (int k = 0; k < 3; k++) d->subtree_linvel[m->body_parentid[i]][k] += d->subtree_linvel[i][k];
    double inv = 1 / fmax(MINVAL, m->body_subtreemass[i]);
    for (int k = 0; k < 3; k++) d->subtree_linvel[i][k] *= inv;
  }
  for (int i = m->nbody - 1; i > 0; i--) {
    int parent = m->body_parentid[i];
    for (int k = 0; k < 3; k++) { dx[k] = d->xipos[i][k] - d->subtree_com[i][k]; dv[k] = body_vel[i][3 + k] - d->subtree_linvel[i][k]; dp[k] = dv[k] * m->body_mass[i]; }
    cross(dL, dx, dp);
    for (int k = 0; k < 3; k++) d->subtree_angmom[i][k] += dL[k];
    for (int k = 0; k < 3; k++) d->subtree_angmom[parent][k] += d->subtree_angmom[i][k];
    for (int k = 0; k < 3; k++) { dx[k] = d->subtree_com[i][k] - d->subtree_com[parent][k]; dv[k] = (d->subtree_linvel[i][k] - d->subtree_linvel[parent][k]) * m->body_subtreemass[i]; }
    cross(dL, dx, dv);
    for (int k = 0; k < 3; k++) d->subtree_angmom[parent][k] += dL[k];
  }
}
static void o_contactForce(const OModel *m, const OData *d, int id, double result[6]) { /* mj_contactForce + mju_decodePyramid [M] */
  (void)m; zero(result, 6);
  const OContact *con = &d->contact[id];
  if (con->efc_address < 0 || con->efc_address >= d->nefc) return;
  const double *f = d->efc_force + con->efc_address;
  if (con->dim == 1) { result[0] = f[0]; return; }
  for (int i = 0; i < 2 * (con->dim - 1); i++) result[0] += f[i];
  for (int i = 1; i < con->dim; i++) result[i] = (f[2 * (i - 1)] - f[2 * (i - 1) + 1]) * con->friction[i - 1];
}
void osim_foot_forces(OSim *c, double cfrc[12]) { /* :1812-1854 */
  const OModel *m = c->m; const OData *d = c->d; double ft[6], fg[3];
  zero(cfrc, 12);
  for (int i = 0; i < d->ncon; i++) {
    int body1 = m->geom_bodyid[d->contact[i].geom1], body2 = m->geom_bodyid[d->contact[i].geom2];
    for (int s = 0; s < 2; s++) {
      int fb = s ? m->right_foot_body : m->left_foot_body;
      if (body1 == fb || body2 == fb) {
        o_contactForce(m, d, i, ft); mulMatTVec3(fg, d->contact[i].frame, ft);
        for (int j = 0; j < 3; j++) cfrc[6 * s + j] += (body1 == fb) ? -fg[j] : fg[j];
      }
    }
  }
}
void osim_heeltoe_forces(OSim *c, double toe_force[6], double heel_force[6]) { /* :1856-1898 */
  const OModel *m = c->m; const OData *d = c->d; double ft[6], fg[3];
  zero(toe_force, 6); zero(heel_force, 6);
  for (int i = 0; i < d->ncon; i++) {
    int body1 = m->geom_bodyid[d->contact[i].geom1], body2 = m->geom_bodyid[d->contact[i].geom2];
    int lf = m->left_foot_body, rf = m->right_foot_body;
    if (body1 == lf || body2 == lf || body1 == rf || body2 == rf) {
      int sign = (body1 == lf || body1 == rf) ? -1 : 1, id = (body1 == rf || body2 == rf) ? 1 : 0, fb = id ? rf : lf;
      o_contactForce(m, d, i, ft); mulMatTVec3(fg, d->contact[i].frame, ft);
      double tw[3], hw[3], v[3]; /* site_xpos of the toe / heel sites (kinematics of the last sub-step) */
      mulMatVec3(v, d->xmat[fb], m->toe_local[id]); for (int k = 0; k < 3; k++) tw[k] = d->xpos[fb][k] + v[k];
      mulMatVec3(v, d->xmat[fb], m->heel_local[id]); for (int k = 0; k < 3; k++) hw[k] = d->xpos[fb][k] + v[k];
      double td[2] = {tw[0] - d->contact[i].pos[0], tw[1] - d->contact[i].pos[1]}, hd[2] = {hw[0] - d->contact[i].pos[0], hw[1] - d->contact[i].pos[1]};
      if (sqrt(td[0] * td[0] + td[1] * td[1]) < sqrt(hd[0] * hd[0] + hd[1] * hd[1])) for (int j = 0; j < 3; j++) toe_force[j + 3 * id] += sign * fg[j];
      else for (int j = 0; j < 3; j++) heel_force[j + 3 * id] += sign * fg[j];
    }
  }
}
void osim_foot_positions(OSim *c, double cpos[6]) { /* :1608-1621 */
  const OModel *m = c->m; const OData *d = c->d;
  copyv(cpos, d->xpos[m->left_foot_body], 3); copyv(cpos + 3, d->xpos[m->right_foot_body], 3);
  double off = sqrt(pow(0.01762, 2) + pow(0.05219, 2));
  cpos[2] -= off; cpos[5] -= off;
}
void osim_foot_velocities(OSim *c, double cvel[12]) { /* :1623-1631: mj_comVel on the CURRENT qvel with the cdof of the last kinematics */
  o_comVel(c->m, c->d);
  copyv(cvel, c->d->cvel[c->m->left_foot_body], 6); copyv(cvel + 6, c->d->cvel[c->m->right_foot_body], 6);
}
void osim_cm_position(OSim *c, double cm_pos[3]) { o_fwdPosition(c->m, c->d); copyv(cm_pos, c->d->subtree_com[0], 3); }          /* :1633-1638 */
void osim_cm_velocity(OSim *c, double cm_vel[3]) { o_fwdPosition(c->m, c->d); o_subtreeVel(c->m, c->d); copyv(cm_vel, c->d->subtree_linvel[0], 3); } /* :1640-1646 */
void osim_angular_momentum(OSim *c, double L[3]) { o_fwdPosition(c->m, c->d); o_subtreeVel(c->m, c->d); copyv(L, c->d->subtree_angmom[0], 3); }       /* :1693-1699 */
int osim_check_obstacle_collision(OSim *c) { /* :1586-1595 */
  for (int i = 0; i < c->d->ncon; i++) if (c->m->geom_user[c->d->contact[i].geom1] == 1 || c->m->geom_user[c->d->contact[i].geom2] == 1) return 1;
  return 0;
}
int osim_check_self_collision(OSim *c) { /* :1597-1606 */
  for (int i = 0; i < c->d->ncon; i++) if (c->m->geom_user[c->d->contact[i].geom1] == 2 && c->m->geom_user[c->d->contact[i].geom2] == 2) return 1;
  return 0;
}
int osim_geom_collision(OSim *c, int geom_group) { /* :1944-1961 */
  for (int i = 0; i < c->d->ncon; i++) {
    int g1 = c->m->geom_group[c->d->contact[i].geom1], g2 = c->m->geom_group[c->d->contact[i].geom2];
    if ((g1 == 1 && g2 == geom_group) || (g2 == 1 && g1 == geom_group)) return 1;
  }
  return 0;
}
void osim_comvel(OSim *c) { o_comVel(c->m, c->d); }

/* ------------------------------------------------------------------ model-constant setters' companion: mj_setConst (src/cassiemujoco.c:949-977)
 * set0 [M]: subtree masses; at qpos0: kinematics, comPos, crb, factorM; body_invweight0 = {trace(Jp inv(M) Jp')/3, trace(Jr inv(M) Jr')/3}
 * with the Jacobians at the body's centre of mass (zero for bodies welded to the world); dof_invweight0 = diag(inv(M)), averaged over the
 * three dofs of a ball joint and over each half of a free joint; setStat [M]: meaninertia = mean diagonal of M(qpos0).
 * The caller's qpos is put back afterwards (cassie_sim_set_const overwrites it anyway). */
void o_setConst(OModel *m, OData *d) {
  int nv = m->nv; double keep[MAXV + 8], jp[3 * MAXV], jr[3 * MAXV], row[MAXV];
  for (int i = 0; i < m->nbody; i++) m->body_subtreemass[i] = m->body_mass[i];
  for (int i = m->nbody - 1; i > 0; i--) m->body_subtreemass[m->body_parentid[i]] += m->body_subtreemass[i];
  copyv(keep, d->qpos, m->nq); copyv(d->qpos, m->qpos0, m->nq);
  o_kinematics(m, d); o_comPos(m, d); o_crb(m, d);
  copyv(d->qLD, d->qM, m->nM); factorI(m, d->qLD, d->qLDiagInv, d->qLDiagSqrtInv);
  for (int b = 1; b < m->nbody; b++) {
    m->body_invweight0[b][0] = m->body_invweight0[b][1] = 0;
    if (!m->body_weldid[b]) continue;
    o_jac(m, d, jp, jr, d->xipos[b], b);
    double tp = 0, tr = 0;
    for (int k = 0; k < 3; k++) {
      copyv(row, jp + k * nv, nv); solveM2(m, d, row); tp += dotn(row, row, nv);
      copyv(row, jr + k * nv, nv); solveM2(m, d, row); tr += dotn(row, row, nv);
    }
    m->body_invweight0[b][0] = tp / 3; m->body_invweight0[b][1] = tr / 3;
  }
  double dg[MAXV];
  for (int i = 0; i < nv; i++) { zero(row, nv); row[i] = 1; solveM2(m, d, row); dg[i] = dotn(row, row, nv); }
  for (int j = 0; j < m->njnt; j++) {
    int da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) { double a = (dg[da] + dg[da + 1] + dg[da + 2]) / 3, b = (dg[da + 3] + dg[da + 4] + dg[da + 5]) / 3; for (int k = 0; k < 3; k++) { m->dof_invweight0[da + k] = a; m->dof_invweight0[da + 3 + k] = b; } }
    else if (m->jnt_type[j] == JNT_BALL) { double a = (dg[da] + dg[da + 1] + dg[da + 2]) / 3; for (int k = 0; k < 3; k++) m->dof_invweight0[da + k] = a; }
    else m->dof_invweight0[da] = dg[da];
  }
  double tr = 0; for (int i = 0; i < nv; i++) tr += d->qM[m->dof_Madr[i]];
  m->meaninertia = tr / nv;
  copyv(d->qpos, keep, m->nq);
}
void osim_just_set_const(OSim *c) { o_setConst(c->m, c->d); }              /* :974-977 */
void osim_set_const(OSim *c) {                                             /* :949-972 */
  static const double qi[28] = {0.0045, 0, 0.4973, 0.9785, -0.0164, 0.01787, -0.2049, -1.1997, 0, 1.4267, 0, -1.5244, 1.5244, -1.5968,
                                -0.0045, 0, 0.4973, 0.9786, 0.00386, -0.01524, -0.2051, -1.1997, 0, 1.4267, 0, -1.5244, 1.5244, -1.5968};
  o_setConst(c->m, c->d);
  { static const double q7[7] = {0, 0, 1.01, 1, 0, 0, 0}; copyv(c->d->qpos, q7, 7); } copyv(c->d->qpos + 7, qi, 28);  /* exactly the 35 constants of :953-958, mju_copy(qpos, qpos_init, 35) at :967: qpos beyond them (cassie_tray_box.xml's cup) keeps its values */
  zero(c->d->qvel, c->m->nv); zero(c->d->qacc, c->m->nv);
  c->d->time = 0;
  o_forward(c->m, c->d);
}
double *osim_model_array(OSim *c, const char *key, int *n) {
  OModel *m = c->m;
#define MARR(name, ptr, cnt) if (!strcmp(key, name)) { *n = (cnt); return (double *)(ptr); }
  MARR("body_mass", m->body_mass, m->nbody) MARR("body_ipos", m->body_ipos, 3 * m->nbody) MARR("dof_damping", m->dof_damping, m->nv)
  MARR("geom_friction", m->geom_friction, 3 * m->ngeom) MARR("body_invweight0", m->body_invweight0, 2 * m->nbody) MARR("dof_invweight0", m->dof_invweight0, m->nv)
  MARR("body_subtreemass", m->body_subtreemass, m->nbody) MARR("meaninertia", &m->meaninertia, 1) MARR("timestep", &m->timestep, 1) MARR("geom_pos", m->geom_pos, 3 * m->ngeom) MARR("geom_quat", m->geom_quat, 4 * m->ngeom) MARR("geom_size", m->geom_size, 3 * m->ngeom) MARR("jnt_stiffness", m->jnt_stiffness, m->njnt) MARR("qpos_spring", m->qpos_spring, m->nq) MARR("body_pos", m->body_pos, 3 * m->nbody)
  *n = 0; return NULL;
}

/* cassie_out_t image from the flat probe layout (oracle/probe_estimator.c): motor pos[10] vel[10], joint pos[6] vel[6], IMU quat[4] gyro[3] accel[3] mag[3] */
void osim_fill_cassie_out(cassie_out_t *o, const double *in) {
  cassie_out_init_(o);
  for (int i = 0; i < 10; i++) { elmo_out_t *e = drive_ptr(o, i); e->position = in[i]; e->velocity = in[10 + i]; }
  for (int i = 0; i < 6; i++) { joint_ptr(o, i)->position = in[20 + i]; joint_ptr(o, i)->velocity = in[26 + i]; }
  for (int k = 0; k < 4; k++) o->pelvis.vectorNav.orientation[k] = in[32 + k];
  for (int k = 0; k < 3; k++) { o->pelvis.vectorNav.angularVelocity[k] = in[36 + k]; o->pelvis.vectorNav.linearAcceleration[k] = in[39 + k]; o->pelvis.vectorNav.magneticField[k] = in[42 + k]; }
}
void osim_set_radio8(cassie_out_t *o, double v) { o->pelvis.radio.channel[8] = v; }

/* ---------------- accessors for the test harness (ctypes) */
OModel *osim_model(OSim *c) { return c->m; }
OData *osim_data(OSim *c) { return c->d; }
cassie_out_t *osim_cassie_out(OSim *c) { return &c->cassie_out; }
int *osim_drive_filter(OSim *c) { return &c->drive_filter[0][0]; }
double *osim_joint_filter_x(OSim *c) { return &c->joint_filter_x[0][0]; }
double *osim_joint_filter_y(OSim *c) { return &c->joint_filter_y[0][0]; }
double *osim_torque_delay(OSim *c) { return &c->torque_delay[0][0]; }
float *osim_hfield_data(OSim *c) { return c->m->hfield_data; }
void osim_forward(OSim *c) { o_forward(c->m, c->d); }
void osim_set_caps(OSim *c, int max_contacts, int max_rows) { c->d->cap_con = max_contacts; c->d->cap_rows = max_rows; }
void osim_mj_step(OSim *c) { o_step(c->m, c->d); }
#define ARR(name, ptr, cnt) if (!strcmp(key, name)) { *n = (cnt); return (double *)(ptr); }
double *osim_array(OSim *c, const char *key, int *n) {
  OData *d = c->d; OModel *m = c->m; int nv = m->nv, nb = m->nbody;
  ARR("qpos", d->qpos, m->nq) ARR("qvel", d->qvel, nv) ARR("qacc", d->qacc, nv) ARR("qacc_warmstart", d->qacc_warmstart, nv)
  ARR("ctrl", d->ctrl, m->nu) ARR("xfrc_applied", d->xfrc_applied, 6 * nb) ARR("qfrc_applied", d->qfrc_applied, nv) ARR("time", &d->time, 1)
  ARR("xpos", d->xpos, 3 * nb) ARR("xquat", d->xquat, 4 * nb) ARR("xmat", d->xmat, 9 * nb) ARR("xipos", d->xipos, 3 * nb)
  ARR("ximat", d->ximat, 9 * nb) ARR("xanchor", d->xanchor, 3 * m->njnt) ARR("xaxis", d->xaxis, 3 * m->njnt)
  ARR("geom_xpos", d->geom_xpos, 3 * m->ngeom) ARR("geom_xmat", d->geom_xmat, 9 * m->ngeom) ARR("subtree_com", d->subtree_com, 3 * nb)
  ARR("cdof", d->cdof, 6 * nv) ARR("cinert", d->cinert, 10 * nb) ARR("crb", d->crb, 10 * nb) ARR("qM", d->qM, m->nM) ARR("qLD", d->qLD, m->nM)
  ARR("qLDiagInv", d->qLDiagInv, nv) ARR("cvel", d->cvel, 6 * nb) ARR("cdof_dot", d->cdof_dot, 6 * nv) ARR("qfrc_bias", d->qfrc_bias, nv)
  ARR("qfrc_passive", d->qfrc_passive, nv) ARR("qfrc_actuator", d->qfrc_actuator, nv) ARR("qfrc_smooth", d->qfrc_smooth, nv)
  ARR("qacc_smooth", d->qacc_smooth, nv) ARR("qfrc_constraint", d->qfrc_constraint, nv) ARR("actuator_velocity", d->actuator_velocity, m->nu)
  ARR("sensordata", d->sensordata, 29) ARR("efc_pos", d->efc_pos, d->nefc) ARR("efc_R", d->efc_R, d->nefc) ARR("efc_D", d->efc_D, d->nefc)
  ARR("efc_aref", d->efc_aref, d->nefc) ARR("efc_b", d->efc_b, d->nefc) ARR("efc_force", d->efc_force, d->nefc) ARR("efc_vel", d->efc_vel, d->nefc)
  ARR("efc_diagApprox", d->efc_diagApprox, d->nefc) ARR("efc_J", d->efc_J, d->nefc * MAXV) ARR("efc_AR", d->efc_AR, d->nefc * d->nefc)
  ARR("cacc", d->cacc, 6 * nb) ARR("subtree_linvel", d->subtree_linvel, 3 * nb) ARR("subtree_angmom", d->subtree_angmom, 3 * nb)
  *n = 0; return NULL;
}
int osim_int(OSim *c, const char *key) {
  OData *d = c->d;
  if (!strcmp(key, "nefc")) return d->nefc; if (!strcmp(key, "ncon")) return d->ncon; if (!strcmp(key, "ne")) return d->ne; if (!strcmp(key, "nl")) return d->nl;
  if (!strcmp(key, "solver_iter")) return d->solver_iter; if (!strcmp(key, "unsupported_pairs")) return d->unsupported_pairs;
  if (!strcmp(key, "dropped_contacts")) return d->dropped_contacts; if (!strcmp(key, "MAXV")) return MAXV;
  if (!strcmp(key, "ngeom")) return c->m->ngeom;
  if (!strcmp(key, "nq")) return c->m->nq; if (!strcmp(key, "nv")) return c->m->nv; if (!strcmp(key, "nbody")) return c->m->nbody;
  return -1;
}
int osim_contact(OSim *c, int i, double *out /* pos3 frame9 dist */, int *geoms) {
  if (i >= c->d->ncon) return 0; OContact *k = &c->d->contact[i];
  copyv(out, k->pos, 3); copyv(out + 3, k->frame, 9); out[12] = k->dist; geoms[0] = k->geom1; geoms[1] = k->geom2; geoms[2] = k->dim; return 1;
}
/* CPU-baseline helper: run `ticks` step_pd ticks on `nsim` private sims in this thread, return 0 */
int osim_run(OSim **sims, int nsim, const pd_in_t *u, int ticks) { state_out_t y; for (int t = 0; t < ticks; t++) for (int i = 0; i < nsim; i++) osim_step_pd(sims[i], u, &y, NULL); return 0; }
