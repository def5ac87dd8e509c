// emu.cpp -- TEST-ONLY host execution of the product's stepper source (cassie-mujoco-sim_b200/csrc/step_core.inl compiled with
// -DCASSIE_EMU: every warp phase becomes a loop over 32 lanes).  It exists so the kernel logic can be diffed against the oracle
// on a machine without a GPU.  It is never linked into, loaded by, or reachable from the product library.
#define CASSIE_EMU 1
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../cassie-mujoco-sim_b200/csrc/devbuild.h"
#include "../../cassie-mujoco-sim_b200/csrc/estimator_host.h"
#include "../../cassie-mujoco-sim_b200/csrc/step_core.inl"

using namespace cassie;

template <typename real> struct Emu {
  HostModel hm; DevModel<real> dm; BuildInfo info; std::vector<real> sm; std::vector<int> ism;
  real qvel[32], qacc_ws[32], xqvel[32], xqacc_ws[32], pd[PD_W], xfrc[XFRC_W], obs[OBS_W], dbg[D_SIZE], cst[CST_W], qM[2 * NM_MAX], aux[AUX_W], cenv[CE_W], task[TASK_W], gait[GAIT_W]; bool use_gait = false; int counters[8]; double est[EST_W]; bool use_est = false; bool use_task = false; bool use_cenv = false, use_ext = true;   // use_ext: run the extended instance (derived-quantity rows on)
  EnvPtrs<real> ptrs() { EnvPtrs<real> E; E.cst = cst; E.dfilt = ism.data(); E.pd = pd; E.xfrc = xfrc; E.obs = obs; E.qM = qM; E.dbg = dbg; E.counters = counters; E.cta_sync = 0; E.nsub = 0; E.task = use_task ? task : nullptr; E.gait = use_gait ? gait : nullptr; E.aux = use_ext ? aux : nullptr; E.cenv = use_cenv ? cenv : nullptr; E.hfield = hfield.empty() ? nullptr : hfield.data(); E.est = use_est ? est : nullptr; E.est_out = obs + OB_EST_OUT; return E; }
  std::vector<float> hfield;
  bool init(const char *path, std::string &err) {
    if (!load_model_any(path, hm, err)) return false;
    if (!build_dev_model(hm, dm, err, &info)) return false;
    sm.assign(scratch_reals_ext(dm.ystride), 0); ism.assign(DFILT_W, 0);
    if (hm.nhfield) hfield.assign((size_t)hm.hfield_nrow[0] * hm.hfield_ncol[0], 0.0f);
    std::vector<real> qpos(QPOS_W_XB), qv(QVEL_W_XB), qa(QVEL_W_XB);
    init_env_rows(hm, qpos.data(), qv.data(), qa.data(), cst, ism.data(), xfrc);
    for (int i = 0; i < 32; i++) { qvel[i] = qv[i]; qacc_ws[i] = qa[i]; xqvel[i] = 0; xqacc_ws[i] = 0; }
    for (int i = 0; i < QPOS_W_XB; i++) sm[S_QPOS + i] = qpos[i];
    std::memset(pd, 0, sizeof pd); std::memset(obs, 0, sizeof obs); std::memset(dbg, 0, sizeof dbg); std::memset(counters, 0, sizeof counters); std::memset(aux, 0, sizeof aux); std::memset(est, 0, sizeof est);
    forward();
    return true;
  }
  void run(int nticks, int mode) { if (use_cenv || use_ext) step_env<real, true, F_ALL>(dm, sm.data(), ptrs(), qvel, qacc_ws, xqvel, xqacc_ws, nticks, mode); else step_env<real, false, F_ALL>(dm, sm.data(), ptrs(), qvel, qacc_ws, xqvel, xqacc_ws, nticks, mode); }
  void forward() { run(1, 1); }
  // per-environment model constants + mj_setConst (mode 3 works at the reference configuration, like the kernel wrapper does)
  void enable_cenv() { if (!use_cenv) { init_cenv_row(dm, cenv, hm, info.geom_dev); use_cenv = true; } }
  int model_set(const char *what, const double *v, int n) {   // same slot mapping as the product's cassie_batch_set_* verbs
    enable_cenv(); int w = 0; cenv_slot(hm, dm.nv, info.geom_dev, what, 0, w);
    if (w < 0 || w != n) return -1;
    for (int i = 0; i < w; i++) { int ww; const int sl = cenv_slot(hm, dm.nv, info.geom_dev, what, i, ww); if (sl >= 0) cenv[sl] = (real)v[i]; }
    return 0;
  }
  void set_const() {
    enable_cenv();
    std::vector<real> keep(sm.begin() + S_QPOS, sm.begin() + S_QPOS + QPOS_W_XB);
    for (int i = 0; i < QPOS_W_XB; i++) sm[S_QPOS + i] = dm.qpos0[i];
    run(1, 3);
    for (int i = 0; i < QPOS_W_XB; i++) sm[S_QPOS + i] = keep[i];
  }
  void query() { run(1, 2); }
  void step(int nticks) { run(nticks, 0); }
};

struct Handle { int fp32; Emu<float> f; Emu<double> d; };

// move / turn / resize a geom (host geom numbering = the reference's) and rebuild the constant block, like cassie_sim_set_geom_name_pos & co.
template <typename E> static int emu_geom_set_(E &e, int g, const double *pos, const double *quat, const double *size) {
  if (g < 0 || g >= e.hm.ngeom) return -1;
  if (pos) for (int k = 0; k < 3; k++) e.hm.geom_pos[3 * g + k] = pos[k];
  if (quat) for (int k = 0; k < 4; k++) e.hm.geom_quat[4 * g + k] = quat[k];
  if (size) for (int k = 0; k < 3; k++) e.hm.geom_size[3 * g + k] = size[k];
  std::string err; return build_dev_model(e.hm, e.dm, err, &e.info) ? 0 : -2;
}


extern "C" {
void *emu_new(const char *path, int fp32) {
  Handle *h = new Handle(); h->fp32 = fp32; std::string err;
  bool ok = fp32 ? h->f.init(path, err) : h->d.init(path, err);
  if (!ok) { fprintf(stderr, "emu: %s\n", err.c_str()); delete h; return nullptr; }
  return h;
}
void emu_free(void *p) { delete (Handle *)p; }
void emu_step(void *p, const double *pd50, int nticks) {
  Handle *h = (Handle *)p;
  if (h->fp32) { for (int i = 0; i < 50; i++) h->f.pd[i] = (float)pd50[i]; h->f.step(nticks); }
  else { for (int i = 0; i < 50; i++) h->d.pd[i] = pd50[i]; h->d.step(nticks); }
}
int emu_set_geom(void *p, int g, const double *pos, const double *quat, const double *size) { Handle *h = (Handle *)p; return h->fp32 ? emu_geom_set_(h->f, g, pos, quat, size) : emu_geom_set_(h->d, g, pos, quat, size); }
int emu_geom_id(void *p, const char *name) { Handle *h = (Handle *)p; return (h->fp32 ? h->f.hm : h->d.hm).geom_id(name); }
void emu_set_hfield(void *p, const float *data, int n) { Handle *h = (Handle *)p; std::vector<float> &dst = h->fp32 ? h->f.hfield : h->d.hfield; for (int i = 0; i < n && i < (int)dst.size(); i++) dst[i] = data[i]; }
int emu_model_set(void *p, const char *what, const double *v, int n) { Handle *h = (Handle *)p; return h->fp32 ? h->f.model_set(what, v, n) : h->d.model_set(what, v, n); }
void emu_set_const(void *p) { Handle *h = (Handle *)p; if (h->fp32) h->f.set_const(); else h->d.set_const(); }
void emu_plain(void *p) { Handle *h = (Handle *)p; if (h->fp32) h->f.use_ext = false; else h->d.use_ext = false; }   // run the plain instance from now on
void emu_enable_cenv(void *p) { Handle *h = (Handle *)p; if (h->fp32) h->f.enable_cenv(); else h->d.enable_cenv(); }
void emu_set_task(void *p, const double *rows60) { Handle *h = (Handle *)p;
  if (h->fp32) { h->f.use_task = rows60 != nullptr; if (rows60) for (int i = 0; i < 60; i++) h->f.task[i] = (float)rows60[i]; }
  else { h->d.use_task = rows60 != nullptr; if (rows60) for (int i = 0; i < 60; i++) h->d.task[i] = rows60[i]; } }
void emu_set_gait(void *p, const double *row21) { Handle *h = (Handle *)p;   // amp[10], phase[10], freq; null: off.  Restarts the tick clock.
  if (h->fp32) { h->f.use_gait = row21 != nullptr; h->f.ism[DF_TICK] = 0; if (row21) for (int i = 0; i < 21; i++) h->f.gait[i] = (float)row21[i]; }
  else { h->d.use_gait = row21 != nullptr; h->d.ism[DF_TICK] = 0; if (row21) for (int i = 0; i < 21; i++) h->d.gait[i] = row21[i]; } }
void emu_enable_est(void *p, int on) { Handle *h = (Handle *)p; if (h->fp32) { h->f.use_est = on; std::memset(h->f.est, 0, sizeof h->f.est); } else { h->d.use_est = on; std::memset(h->d.est, 0, sizeof h->d.est); } }   // (re)start the in-kernel estimator
void emu_query(void *p) { Handle *h = (Handle *)p; if (h->fp32) h->f.query(); else h->d.query(); }
void emu_forward(void *p) { Handle *h = (Handle *)p; if (h->fp32) h->f.forward(); else h->d.forward(); }
// generic get/set of named state as doubles.  names: qpos qvel qacc_ws cst obs dbg xfrc ; ints: dfilt counters
int emu_get(void *p, const char *name, double *out, int n) {
  Handle *h = (Handle *)p; std::string k(name);
#define GET(T, E) { const T *src = nullptr; int cnt = 0; \
  if (k == "qpos") { src = E.sm.data() + S_QPOS; cnt = QPOS_W_XB; } else if (k == "xqvel") { src = E.xqvel; cnt = 6; } else if (k == "qvel") { src = E.qvel; cnt = 32; } else if (k == "qacc_ws") { src = E.qacc_ws; cnt = 32; } \
  else if (k == "cst") { src = E.cst; cnt = CST_W; } else if (k == "obs") { src = E.obs; cnt = OBS_W; } else if (k == "dbg") { src = E.dbg; cnt = D_SIZE; } \
  else if (k == "xfrc") { src = E.xfrc; cnt = XFRC_W; } else if (k == "aux") { src = E.aux; cnt = AUX_W; } else if (k == "cenv") { src = E.cenv; cnt = CE_W; } else if (k == "est_out") { src = E.obs + OB_EST_OUT; cnt = EO_W; } \
  if (src) { if (cnt > n) cnt = n; for (int i = 0; i < cnt; i++) out[i] = (double)src[i]; return cnt; } \
  if (k == "dfilt") { int c2 = DFILT_W < n ? DFILT_W : n; for (int i = 0; i < c2; i++) out[i] = E.ism[i]; return c2; } \
  if (k == "counters") { int c2 = 8 < n ? 8 : n; for (int i = 0; i < c2; i++) out[i] = E.counters[i]; return c2; } }
  if (h->fp32) GET(float, h->f) else GET(double, h->d)
  return -1;
}
int emu_set(void *p, const char *name, const double *in, int n) {
  Handle *h = (Handle *)p; std::string k(name);
#define SET(T, E) { T *dst = nullptr; int cnt = 0; \
  if (k == "qpos") { dst = E.sm.data() + S_QPOS; cnt = QPOS_W_XB; } else if (k == "xqvel") { dst = E.xqvel; cnt = 6; } else if (k == "qvel") { dst = E.qvel; cnt = 32; } else if (k == "qacc_ws") { dst = E.qacc_ws; cnt = 32; } \
  else if (k == "cst") { dst = E.cst; cnt = CST_W; } else if (k == "xfrc") { dst = E.xfrc; cnt = XFRC_W; } else if (k == "cenv") { dst = E.cenv; cnt = CE_W; } \
  if (dst) { if (cnt > n) cnt = n; for (int i = 0; i < cnt; i++) dst[i] = (T)in[i]; return cnt; } }
  if (h->fp32) SET(float, h->f) else SET(double, h->d)
  return -1;
}
}
