// step_core.inl -- the batched Cassie stepper: one WARP per environment, lanes = dofs / bodies / constraint rows.
//
// This file is the hot path named by BASELINE.json: everything cassie_sim_step_pd does for one environment
// (/root/reference/src/cassiemujoco.c:1147-1157 -> :1137-1145 -> :1115-1135), i.e.
//   pd_input_step (motor-PD branch)  ->  cassie_core_sim_step (safety layer)  ->  motor model + 6-tick torque delay (:638-664)
//   ->  encoder / filter emulation (:558-635, 737-774)  ->  mj_step1 + mj_step2 (MuJoCo 2.1.0 semantics: kinematics, comPos, CRB,
//   sparse L'DL, collision, constraint assembly, PGS with warm start, sensors, implicit-damping Euler)
// fused so that a launch advances every environment `nticks` control ticks with no host round trip.
//
// The code is written as warp-synchronous PHASES: `LANES { ... } ENDL` runs the body once per lane with lane index `l`, and
// all cross-lane traffic happens between phases (shared-memory scratch `sm`, warp shuffles via ALLSUM / BCAST / EXSCAN_INT).
// On the GPU (kernels.cu) a phase is straight-line code followed by __syncwarp().  With -DCASSIE_EMU (tests/emu only, never
// linked into the product) a phase is a `for (l = 0..31)` loop and lane variables are arrays, which lets the exact same
// source be executed and diffed against the oracle on a machine without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include "devmodel.h"
#include "cassie_tree_gen.inc"

#ifdef CASSIE_EMU
#define CFN inline
#define CNOINLINE inline
#define DECL_LANE
#define STAGE_SYNC(on)
#define LANES for (int l = 0; l < 32; ++l) {
#define ENDL }
#define LANES_NS for (int l = 0; l < 32; ++l) {
#define ENDL_NS }
#define LV(T, name) T name[32]
#define L(name) name[l]
#define LP(T, name) T(&name)[32]
#define LVA(T, name, N) T name[N][32]
#define LA(name, i) name[i][l]
#define LANE0(v) v[0]
#define ALLSUM(v) do { auto s_ = v[0]; for (int i_ = 1; i_ < 32; ++i_) s_ += v[i_]; for (int i_ = 0; i_ < 32; ++i_) v[i_] = s_; } while (0)
#define ALLMAX(v) do { auto s_ = v[0]; for (int i_ = 1; i_ < 32; ++i_) s_ = s_ > v[i_] ? s_ : v[i_]; for (int i_ = 0; i_ < 32; ++i_) v[i_] = s_; } while (0)
#define BCAST(dst, src, lane) do { auto s_ = src[lane]; for (int i_ = 0; i_ < 32; ++i_) dst[i_] = s_; } while (0)
#define EXSCAN_INT(v, total) do { int a_ = 0; for (int i_ = 0; i_ < 32; ++i_) { int t_ = v[i_]; v[i_] = a_; a_ += t_; } total = a_; } while (0)
#define SHFLV(dst, src, idx) do { auto t0_ = src[0]; decltype(t0_) t_[32]; for (int l = 0; l < 32; ++l) t_[l] = src[(idx)]; for (int l = 0; l < 32; ++l) dst[l] = t_[l]; } while (0)
#define BALLOT(mask, v) do { uint32_t m_ = 0; for (int i_ = 0; i_ < 32; ++i_) if (v[i_]) m_ |= 1u << i_; mask = m_; } while (0)
#else
#define CFN __device__ __forceinline__
#define CNOINLINE __device__ __noinline__   // cold / register-hungry stages: their own register allocation, no pressure on the stepping loop
#define DECL_LANE const int l = threadIdx.x & 31;
// optional CTA-wide rendezvous between stages: the warps of a CTA (one environment each) then walk the code together and share
// instruction-cache lines; `on` is uniform over the CTA
#define STAGE_SYNC(on) do { if (on) __syncthreads(); } while (0)
#define LANES {
#define ENDL } __syncwarp();
#define LANES_NS {
#define ENDL_NS }
#define LV(T, name) T name
#define L(name) name
#define LP(T, name) T &name
#define LVA(T, name, N) T name[N]
#define LA(name, i) name[i]
#define LANE0(v) v
#define ALLSUM(v) do { for (int o_ = 16; o_ > 0; o_ >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o_); } while (0)
#define ALLMAX(v) do { for (int o_ = 16; o_ > 0; o_ >>= 1) v = mmax(v, __shfl_xor_sync(0xffffffffu, v, o_)); } while (0)
#define BCAST(dst, src, lane) dst = __shfl_sync(0xffffffffu, src, lane)
#define SHFLV(dst, src, idx) dst = __shfl_sync(0xffffffffu, src, idx)
#define BALLOT(mask, v) mask = __ballot_sync(0xffffffffu, v)
#define EXSCAN_INT(v, total) do { int x_ = v; for (int o_ = 1; o_ < 32; o_ <<= 1) { int y_ = __shfl_up_sync(0xffffffffu, x_, o_); if (l >= o_) x_ += y_; } total = __shfl_sync(0xffffffffu, x_, 31); v = x_ - v; } while (0)
#endif

namespace cassie {

// per-environment rows that stay in global memory (L1/L2 resident; only the owning warp touches them)
template <typename real> struct EnvPtrs {
  real *cst;          // [CST_W] controller / sensor state
  int *dfilt;         // [DFILT_W] drive FIR taps
  const real *pd;     // [PD_W] motor-PD row held for the launch
  const real *xfrc;   // [XFRC_W]
  const real *task;   // [TASK_W] task-space PD rows (pd_in_t taskPd of both legs) or null
  const real *gait;   // [GAIT_W] open-loop gait on the motor-PD targets: pTarget_i(t) = pd.pTarget_i + amp_i sin(2 pi f t + phase_i), t = ticks since reset x 0.5 ms; or null
  real *obs;          // [OBS_W] or null
  real *qM;           // [2 NM_MAX] scratch: M (debug dump / set_const only), then the factor of M + h B carried from the CRB stage to the Euler stage
  const float *hfield; // [nrow*ncol] normalised elevations of this env's terrain, or null
  real *dbg;          // [D_SIZE] or null
  real *aux;          // [AUX_W] derived-quantity row or null
  real *cenv;         // [CE_W] per-environment model constants (domain randomisation) or null: then the shared model block's values apply
  double *est;        // [EST_W] in-kernel estimator state or null (extended instance only)
  real *est_out;      // [EO_W] its outputs
  int *counters;      // [8]
  int cta_sync;       // 1: the CTA's warps rendezvous at the stage boundaries (STAGE_SYNC); only the step / forward modes
  int nsub;           // > 0: physics sub-steps per control tick for this launch instead of round(5e-4 / timestep) (cassie_sim_step_pd_no2khz: 1)
};

// ------------------------------------------------------------------ scalar math on float / double
CFN float msqrt(float x) { return sqrtf(x); }
CFN double msqrt(double x) { return sqrt(x); }
CFN float matan2(float y, float x) { return atan2f(y, x); }
CFN double matan2(double y, double x) { return atan2(y, x); }
CFN float masin(float x) { return asinf(x); }
CFN double masin(double x) { return asin(x); }
CFN int mmax(int a, int b) { return a > b ? a : b; }
CFN float mabs(float x) { return fabsf(x); }
CFN double mabs(double x) { return fabs(x); }
CFN float mmax(float a, float b) { return fmaxf(a, b); }
CFN double mmax(double a, double b) { return fmax(a, b); }
CFN float mmin(float a, float b) { return fminf(a, b); }
CFN double mmin(double a, double b) { return fmin(a, b); }
CFN float mpow(float a, float b) { return powf(a, b); }
CFN double mpow(double a, double b) { return pow(a, b); }
CFN float msin(float x) { return sinf(x); }
CFN double msin(double x) { return sin(x); }
CFN void msincos(float x, float *s, float *c) { sincosf(x, s, c); }
CFN void msincos(double x, double *s, double *c) { sincos(x, s, c); }
CFN float mrcp(float x) {
#ifdef CASSIE_EMU
  return 1.0f / x;
#else
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;   // one MUFU.RCP (1 ulp) instead of the 10-instruction correctly rounded reciprocal
#endif
}
CFN double mrcp(double x) { return 1.0 / x; }
template <typename real> CFN real minval();
template <> CFN float minval<float>() { return 1e-15f; }
template <> CFN double minval<double>() { return 1e-15; }

template <typename real> CFN real dot3(const real *a, const real *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename real> CFN void cross3(real *r, const real *a, const real *b) {
  real t0 = a[1] * b[2] - a[2] * b[1], t1 = a[2] * b[0] - a[0] * b[2], t2 = a[0] * b[1] - a[1] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2;
}
template <typename real> CFN real normalize3(real *v) {
  real n = msqrt(dot3(v, v));
  if (n < minval<real>()) { v[0] = 1; v[1] = 0; v[2] = 0; return 0; }
  real inv = real(1) / n; v[0] *= inv; v[1] *= inv; v[2] *= inv; return n;
}
template <typename real> CFN void normalize4(real *q) {
  real n = msqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < minval<real>()) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  real inv = real(1) / n; q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
template <typename real> CFN void mul_quat(real *r, const real *a, const real *b) {
  real t0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], t1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real t2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], t3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}
template <typename real> CFN void quat2mat(real *m, const real *q) {
  real q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3], q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3],
       q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02); m[3] = 2 * (q12 + q03); m[5] = 2 * (q23 - q01); m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
template <typename real> CFN void mat_vec(real *r, const real *m, const real *v) {
  real t0 = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], t1 = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], t2 = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = t0; r[1] = t1; r[2] = t2;
}
template <typename real> CFN void matT_vec(real *r, const real *m, const real *v) {
  real t0 = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], t1 = m[1] * v[0] + m[4] * v[1] + m[7] * v[2], t2 = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = t0; r[1] = t1; r[2] = t2;
}
// 10-number spatial inertia times (angular; linear) motion vector, MuJoCo's c-frame convention
template <typename real> CFN void mul_inert_vec(real *r, const real *i, const real *v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
template <typename real> CFN void cross_motion(real *r, const real *vel, const real *v) {
  real t[3]; cross3(r, vel, v); cross3(r + 3, vel, v + 3); cross3(t, vel + 3, v);
  r[3] += t[0]; r[4] += t[1]; r[5] += t[2];
}
template <typename real> CFN void cross_force(real *r, const real *vel, const real *f) {
  real t[3]; cross3(r, vel, f); cross3(t, vel + 3, f + 3);
  r[0] += t[0]; r[1] += t[1]; r[2] += t[2];
  cross3(r + 3, vel, f + 3);
}
template <typename real> CFN real clampr(real x, real lo, real hi) { return mmin(mmax(x, lo), hi); }
// MuJoCo's impedance sigmoid (solimp = dmin dmax width midpoint power)
template <typename real> CFN real impedance(const real *solimp, real pos, real margin) {
  if (solimp[0] == solimp[1] || solimp[2] <= minval<real>()) return real(0.5) * (solimp[0] + solimp[1]);
  real x = mabs((pos - margin) / solimp[2]);
  if (x >= 1) return solimp[1];
  if (x <= 0) return solimp[0];
  real y, p = solimp[4], mid = solimp[3];
  if (p == 1) y = x;
  else if (p == 2) y = (x <= mid) ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid);
  else if (x <= mid) y = mpow(x, p) / mpow(mid, p - 1);
  else y = 1 - mpow(1 - x, p) / mpow(1 - mid, p - 1);
  return solimp[0] + y * (solimp[1] - solimp[0]);
}
// contact frame from a normal and an optional tangent hint (zero hint -> canonical choice)
template <typename real> CFN void make_frame(real *f) {
  normalize3(f);
  if (msqrt(dot3(f + 3, f + 3)) < real(0.5)) { f[3] = f[4] = f[5] = 0; if (f[1] < real(0.5) && f[1] > real(-0.5)) f[4] = 1; else f[5] = 1; }
  real s = dot3(f, f + 3);
  f[3] -= f[0] * s; f[4] -= f[1] * s; f[5] -= f[2] * s;
  normalize3(f + 3); cross3(f + 6, f, f + 3);
}

// ------------------------------------------------------------------ sparse L'DL on the dof tree (lane = dof where it matters)
// in-place factorisation of sm[S_QLD..] (already holding M); writes 1/D and 1/sqrt(D)
// q2 (optional): a second matrix of the same sparsity (M + h B for the implicit-damping Euler step) factored in the same pass: the schedule
// decode and the loop overhead are shared; every entry sees exactly the operations a separate factorisation would apply
// four consecutive reals of a 16-byte aligned shared-memory record in one (fp32) or two (fp64) 128-bit loads
template <typename real> struct Row4 { real x, y, z, w; };
template <typename real> CFN Row4<real> load_row4(const real *p) {
#ifdef CASSIE_EMU
  return Row4<real>{p[0], p[1], p[2], p[3]};
#else
  Row4<real> r;
  if (sizeof(real) == 4) { const float4 v = *reinterpret_cast<const float4 *>(p); r.x = (real)v.x; r.y = (real)v.y; r.z = (real)v.z; r.w = (real)v.w; }
  else { const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2); r.x = (real)a.x; r.y = (real)a.y; r.z = (real)b.x; r.w = (real)b.y; }
  return r;
#endif
}
template <typename real> CFN void factor_ld(const DevModel<real> &cm, real *sm, real *q2 = (real *)0) {
  DECL_LANE
  real *qLD = sm + S_QLD, *dinv2 = sm + S_VEC + 128;
  if (cm.nfac > 0) {
    // Balanced schedule: eliminating dof k means dk(dk+1)/2 independent updates  M(anc_t, .)[c] -= M(k, .)[t + c] * f_t,  f_t = M(k, anc_t) / D_k;  a
    // precomputed table deals them round-robin to the lanes, two per lane and trip so that their loads overlap.  f_t is formed on the fly from the
    // still un-normalised row k, which is never written again: the rows are normalised (L = M(k, .) / D_k) in one parallel pass after the last step.
    // One warp rendezvous per step (the next dof's row must have seen this step's updates).
    for (int k = cm.nv - 1; k >= 0; --k) {
      if (cm.dof_depth[k] == 0) continue;
      const int kk = cm.dof_Madr[k], p0 = cm.fac_start[k], p1 = cm.fac_start[k + 1];
      const real rinv = mrcp(qLD[kk]), rinv2 = q2 ? mrcp(q2[kk]) : real(0);
      if (p1 - p0 <= 32) {   // the shallow dofs: one update per lane
        LANES
          if (p0 + l < p1) {
            const uint32_t ea = cm.fac_pairs[p0 + l]; const int da = ea & 0xfffu, sa = (ea >> 12) & 0xfffu, ta = kk + (int)(ea >> 24);
            if (q2) q2[da] -= q2[sa] * (q2[ta] * rinv2);
            qLD[da] -= qLD[sa] * (qLD[ta] * rinv);
          }
        ENDL
        continue;
      }
      LANES
        for (int p = p0 + l; p < p1; p += 64) {
          const bool two = p + 32 < p1;
          const uint32_t ea = cm.fac_pairs[p], eb = two ? cm.fac_pairs[p + 32] : ea;
          const int da = ea & 0xfffu, sa = (ea >> 12) & 0xfffu, ta = kk + (int)(ea >> 24), db = eb & 0xfffu, sb = (eb >> 12) & 0xfffu, tb = kk + (int)(eb >> 24);
          const real va = qLD[da] - qLD[sa] * (qLD[ta] * rinv), vb = qLD[db] - qLD[sb] * (qLD[tb] * rinv);
          if (q2) {
            const real wa = q2[da] - q2[sa] * (q2[ta] * rinv2), wb = q2[db] - q2[sb] * (q2[tb] * rinv2);
            q2[da] = wa; if (two) q2[db] = wb;
          }
          qLD[da] = va; if (two) qLD[db] = vb;
        }
      ENDL
    }
    LANES if (l < cm.nv) { const real d = qLD[cm.dof_Madr[l]]; sm[S_DINV + l] = mrcp(d); sm[S_DSQI + l] = mrcp(msqrt(d)); if (q2) dinv2[l] = mrcp(q2[cm.dof_Madr[l]]); } ENDL
    LANES
      for (int p = l; p < cm.ntri; p += 32) { const uint32_t e = cm.tri[p]; const int i = e >> 24, a = e & 0xffffu; qLD[a] *= sm[S_DINV + i]; if (q2) q2[a] *= dinv2[i]; }
    ENDL
    return;
  }
  LV(real, tmp);
  for (int k = cm.nv - 1; k >= 0; --k) {
    const int dk = cm.dof_depth[k];
    if (dk == 0) continue;
    const int kk = cm.dof_Madr[k];
    const uint32_t anc = cm.dof_ancmask[k];
    for (int pass = 0; pass < (q2 ? 2 : 1); ++pass) {
      real *Q = pass ? q2 : qLD;
      LANES  // lane i (an ancestor of k): row_i -= row_k[t..] * (M(k,i)/M(k,k)),  t = depth(k) - depth(i)
        L(tmp) = 0;
        if ((anc >> l) & 1u) {
          const int di = cm.dof_depth[l], t = dk - di, ia = cm.dof_Madr[l];
          const real f = Q[kk + t] / Q[kk];
          const real *rk = Q + kk + t; real *ri = Q + ia;
          for (int c = 0; c <= di; ++c) ri[c] -= rk[c] * f;
          L(tmp) = f;
        }
      ENDL
      LANES if ((anc >> l) & 1u) Q[kk + dk - cm.dof_depth[l]] = L(tmp); ENDL
    }
  }
  LANES if (l < cm.nv) { const real d = qLD[cm.dof_Madr[l]]; sm[S_DINV + l] = mrcp(d); sm[S_DSQI + l] = mrcp(msqrt(d)); } ENDL
}
// triangular sweeps on the factorisation in sm[S_QLD..] (one value per lane = dof); each is nv broadcast steps
// x <- inv(L') x
template <typename real> CFN void sweep_lt(const DevModel<real> &cm, const real *sm, LP(real, x)) {
  DECL_LANE
  LV(real, xi); LV(int, dl);
  LANES_NS L(dl) = (l < cm.nv) ? cm.dof_depth[l] : 0; ENDL_NS
  for (int i = cm.nv - 1; i >= 0; --i) {
    if (cm.dof_depth[i] == 0) continue;
    BCAST(xi, x, i);
    const uint32_t anc = cm.dof_ancmask[i]; const int base = S_QLD + cm.dof_Mrow[i];
    LANES_NS if ((anc >> l) & 1u) L(x) -= sm[base - L(dl)] * L(xi); ENDL_NS
  }
}
// x <- inv(L) x
template <typename real> CFN void sweep_l(const DevModel<real> &cm, const real *sm, LP(real, x)) {
  DECL_LANE
  LV(real, xi); LV(int, ptr); LV(uint32_t, ancl);
  // lane l walks its row of L from the root-most ancestor entry towards the diagonal: ancestors arrive in increasing dof order
  LANES_NS L(ptr) = (l < cm.nv) ? S_QLD + cm.dof_Mrow[l] : 0; L(ancl) = (l < cm.nv) ? cm.dof_ancmask[l] : 0u; ENDL_NS
  for (int j = 0; j < cm.nv; ++j) {
    BCAST(xi, x, j);
    LANES_NS if ((L(ancl) >> j) & 1u) { L(x) -= sm[L(ptr)] * L(xi); L(ptr) -= 1; } ENDL_NS
  }
}
// the same two sweeps for the tree csrc/cassie_tree_gen.inc was generated for: straight-line steps with the two mirrored legs walked side by side
// (19 dependent steps instead of 31: leg lanes take their own leg's value through a per-lane shuffle source, base lanes take both)
#define LNAME_(v) L(v)
template <typename real> CFN void sweep_lt_spec(const DevModel<real> &cm, const real *sm, LP(real, x)) {
  DECL_LANE
  LV(real, xi); LV(real, xj); LV(uint32_t, qbit); LV(int, lsel); LV(int, po);
  LANES_NS { const int leg2 = l >= 19 ? 1 : 0; L(lsel) = leg2 ? 13 : 0; L(qbit) = l < 6 ? 0x80000000u : 1u << (l - 6 - L(lsel)); L(po) = S_QLD - cm.dof_depth[l] + (leg2 ? CT19_SYM_MADR : 0); } ENDL_NS
  CT19_SWEEP_LT(x, xi, xj, SHFLV, BCAST, LANES_NS, ENDL_NS, LNAME_, sm, L(qbit), L(lsel), L(po), S_QLD);
}
template <typename real> CFN void sweep_l_spec(const DevModel<real> &cm, const real *sm, LP(real, x)) {
  DECL_LANE
  LV(real, xi); LV(uint32_t, qbit); LV(int, lsel); LV(int, pr);
  LANES_NS { const int leg2 = l >= 19 ? 1 : 0; L(lsel) = leg2 ? 13 : 0; L(qbit) = l < 6 ? 0u : 1u << (l - 6 - L(lsel)); L(pr) = S_QLD + cm.dof_Mrow[l]; } ENDL_NS
  CT19_SWEEP_L(x, xi, SHFLV, BCAST, LANES_NS, ENDL_NS, LNAME_, sm, L(qbit), L(lsel), L(pr));
}
// x <- inv(L'DL) x
#if defined(CASSIE_EMU) || !defined(CASSIE_SHARED_SOLVES)
template <typename real> CFN void solve_m(const DevModel<real> &cm, const real *sm, LP(real, x)) {
  DECL_LANE
  if (cm.spec19) sweep_lt_spec(cm, sm, x); else sweep_lt(cm, sm, x);
  LANES_NS if (l < cm.nv) L(x) *= sm[S_DINV + l]; ENDL_NS
  if (cm.spec19) sweep_l_spec(cm, sm, x); else sweep_l(cm, sm, x);
}
template <typename real> CFN void solve_l(const DevModel<real> &cm, const real *sm, LP(real, x)) { if (cm.spec19) sweep_l_spec(cm, sm, x); else sweep_l(cm, sm, x); }
#else
// -DCASSIE_SHARED_SOLVES: the three solves of a sub-step (qacc_smooth, the constraint correction, the implicit-damping Euler step) share ONE copy of the
// sweeps' code (-1.6 k static instructions; the hot footprint is larger than the instruction cache).  Measured A/B on config 2: 25.6 vs 25.9 M with the
// inlined copies, same end to end -- so the inlined copies are the default.
// l_only: x <- inv(L) x (the second sweep alone)
template <typename real> CNOINLINE real solve_shared(const DevModel<real> &cm, const real *sm, real x, int l_only) {
  DECL_LANE
  if (!l_only) {
    if (cm.spec19) sweep_lt_spec(cm, sm, x); else sweep_lt(cm, sm, x);
    if (l < cm.nv) x *= sm[S_DINV + l];
  }
  if (cm.spec19) sweep_l_spec(cm, sm, x); else sweep_l(cm, sm, x);
  return x;
}
template <typename real> CFN void solve_m(const DevModel<real> &cm, const real *sm, real &x) { x = solve_shared(cm, sm, x, 0); }
template <typename real> CFN void solve_l(const DevModel<real> &cm, const real *sm, real &x) { x = solve_shared(cm, sm, x, 1); }
#endif

// closest point on triangle abc to p (Ericson, Real-Time Collision Detection 5.1.5); true when it lies strictly inside the face
template <typename real> CFN bool closest_pt_tri(const real *p, const real *a, const real *b, const real *c, real *q) {
  real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]}, ap[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
  const real d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0 && d2 <= 0) { q[0] = a[0]; q[1] = a[1]; q[2] = a[2]; return false; }
  real bp[3] = {p[0] - b[0], p[1] - b[1], p[2] - b[2]};
  const real d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0 && d4 <= d3) { q[0] = b[0]; q[1] = b[1]; q[2] = b[2]; return false; }
  const real vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { const real v = d1 / (d1 - d3); q[0] = a[0] + v * ab[0]; q[1] = a[1] + v * ab[1]; q[2] = a[2] + v * ab[2]; return false; }
  real cp[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
  const real d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0 && d5 <= d6) { q[0] = c[0]; q[1] = c[1]; q[2] = c[2]; return false; }
  const real vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { const real w = d2 / (d2 - d6); q[0] = a[0] + w * ac[0]; q[1] = a[1] + w * ac[1]; q[2] = a[2] + w * ac[2]; return false; }
  const real va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { const real w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); q[0] = b[0] + w * (c[0] - b[0]); q[1] = b[1] + w * (c[1] - b[1]); q[2] = b[2] + w * (c[2] - b[2]); return false; }
  const real den = real(1) / (va + vb + vc), v = vb * den, w = vc * den;
  q[0] = a[0] + ab[0] * v + ac[0] * w; q[1] = a[1] + ab[1] * v + ac[1] * w; q[2] = a[2] + ab[2] * v + ac[2] * w;
  return true;
}
// height field vs sphere: the deepest analytic contact against the triangulated surface under the sphere (own definition);
// hf = normalised elevations [nrow][ncol] in global memory; the hfield frame is axis aligned at hpos
template <typename real> CFN bool hfield_sphere(const DevModel<real> &cm, const float *hf, const real *hpos, const real *sp, real r, real margin, real *dist, real *nrm) {
  const real sx = cm.hf_size[0], sy = cm.hf_size[1], sz = cm.hf_size[2]; const int nrow = cm.hf_nrow, ncol = cm.hf_ncol;
  real pl[3] = {sp[0] - hpos[0], sp[1] - hpos[1], sp[2] - hpos[2]};
  if (mabs(pl[0]) > sx + r || mabs(pl[1]) > sy + r || pl[2] - r > sz + margin) return false;
  const real dx = 2 * sx / (ncol - 1), dy = 2 * sy / (nrow - 1);
  int c0 = (int)floor((pl[0] - r + sx) / dx), c1 = (int)floor((pl[0] + r + sx) / dx), r0 = (int)floor((pl[1] - r + sy) / dy), r1 = (int)floor((pl[1] + r + sy) / dy);
  c0 = c0 < 0 ? 0 : (c0 > ncol - 2 ? ncol - 2 : c0); c1 = c1 < 0 ? 0 : (c1 > ncol - 2 ? ncol - 2 : c1);
  r0 = r0 < 0 ? 0 : (r0 > nrow - 2 ? nrow - 2 : r0); r1 = r1 < 0 ? 0 : (r1 > nrow - 2 ? nrow - 2 : r1);
  real best = real(1e30), bn0 = 0, bn1 = 0, bn2 = 1;
  for (int rr = r0; rr <= r1; ++rr) for (int cc = c0; cc <= c1; ++cc) {
    const real x0 = -sx + cc * dx, y0 = -sy + rr * dy;
    const real h00 = (real)hf[rr * ncol + cc] * sz, h10 = (real)hf[rr * ncol + cc + 1] * sz, h01 = (real)hf[(rr + 1) * ncol + cc] * sz, h11 = (real)hf[(rr + 1) * ncol + cc + 1] * sz;
    const real v00[3] = {x0, y0, h00}, v10[3] = {x0 + dx, y0, h10}, v01[3] = {x0, y0 + dy, h01}, v11[3] = {x0 + dx, y0 + dy, h11};
    for (int t = 0; t < 2; ++t) {
      const real *a = t ? v10 : v00, *b = t ? v11 : v10, *c = v01;
      real e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]}, n[3], q[3], dd, nn[3];
      cross3(n, e1, e2); normalize3(n);
      if (closest_pt_tri(pl, a, b, c, q)) { real d[3] = {pl[0] - a[0], pl[1] - a[1], pl[2] - a[2]}; dd = dot3(n, d) - r; nn[0] = n[0]; nn[1] = n[1]; nn[2] = n[2]; }
      else {
        real v[3] = {pl[0] - q[0], pl[1] - q[1], pl[2] - q[2]}; const real len = msqrt(dot3(v, v));
        if (dot3(v, n) < 0 || len < real(1e-12)) continue;
        dd = len - r; nn[0] = v[0] / len; nn[1] = v[1] / len; nn[2] = v[2] / len;
      }
      if (dd < margin && dd < best) { best = dd; bn0 = nn[0]; bn1 = nn[1]; bn2 = nn[2]; }
    }
  }
  if (best > real(1e29)) return false;
  *dist = best; nrm[0] = bn0; nrm[1] = bn1; nrm[2] = bn2;
  return true;
}

// ---- box primitives (cassie_tray_box.xml): analytic definitions of our own; the CPU checker restates the same rules.
// geom record g = [pos3, z-axis3, x-axis3, y-axis3] (world); local -> world: v = x*vx + y*vy + z*vz
template <typename real> CFN void box_to_world(const real *g, const real *l3, real *w) {
  w[0] = g[6] * l3[0] + g[9] * l3[1] + g[3] * l3[2]; w[1] = g[7] * l3[0] + g[10] * l3[1] + g[4] * l3[2]; w[2] = g[8] * l3[0] + g[11] * l3[1] + g[5] * l3[2];
}
template <typename real> CFN void box_to_local(const real *g, const real *w, real *l3) {
  l3[0] = g[6] * w[0] + g[7] * w[1] + g[8] * w[2]; l3[1] = g[9] * w[0] + g[10] * w[1] + g[11] * w[2]; l3[2] = g[3] * w[0] + g[4] * w[1] + g[5] * w[2];
}
template <typename real> CFN void box_corner(const real *g, const real *s, int i, real *out) {
  real l3[3] = {(i & 1) ? s[0] : -s[0], (i & 2) ? s[1] : -s[1], (i & 4) ? s[2] : -s[2]}, w[3];
  box_to_world(g, l3, w); out[0] = g[0] + w[0]; out[1] = g[1] + w[1]; out[2] = g[2] + w[2];
}
// sphere (geom1) vs box (geom2): 0 or 1 contact; normal from the sphere to the box
template <typename real> CFN int sphere_box(real margin, const real *sc, real r, const real *gb, const real *s, real *pos, real *nrm, real *dist) {
  real t[3] = {sc[0] - gb[0], sc[1] - gb[1], sc[2] - gb[2]}, cl[3], q[3], nl[3], dd; bool inside = true;
  box_to_local(gb, t, cl);
  for (int k = 0; k < 3; ++k) { q[k] = clampr(cl[k], -s[k], s[k]); if (q[k] != cl[k]) inside = false; }
  if (inside) {
    int ax = 0; real best = real(1e30);
    for (int k = 0; k < 3; ++k) { const real pen = s[k] - mabs(cl[k]); if (pen < best) { best = pen; ax = k; } }
    nl[0] = nl[1] = nl[2] = 0; nl[ax] = cl[ax] >= 0 ? real(1) : real(-1); q[ax] = nl[ax] * s[ax]; dd = -best - r;
  } else {
    real dv[3] = {cl[0] - q[0], cl[1] - q[1], cl[2] - q[2]}; const real len = msqrt(dot3(dv, dv));
    if (len - r >= margin) return 0;
    nl[0] = dv[0] / len; nl[1] = dv[1] / len; nl[2] = dv[2] / len; dd = len - r;
  }
  if (dd >= margin) return 0;
  real nw[3], qw[3]; box_to_world(gb, nl, nw); box_to_world(gb, q, qw);
  *dist = dd;
  for (int k = 0; k < 3; ++k) { nrm[k] = -nw[k]; pos[k] = gb[k] + qw[k] + nw[k] * dd * real(0.5); }
  return 1;
}
// all box pair kinds; g1 / g2 = geom records, s1 / s2 = geom sizes; returns the contact count (<= 4)
template <typename real> CNOINLINE int box_pair(int kind, real margin, const real *g1, const real *g2, const real *s1, const real *s2, real (*cp)[3], real (*cn)[3], real *cdst) {
  int cnt = 0;
  if (kind == PAIR_PLANE_BOX) {            // every corner below the plane, at most 4, in corner order
    const real *n = g1 + 3;
    for (int i = 0; i < 8 && cnt < 4; ++i) {
      real v[3]; box_corner(g2, s2, i, v);
      real t[3] = {v[0] - g1[0], v[1] - g1[1], v[2] - g1[2]}; const real dd = dot3(t, n);
      if (dd >= margin) continue;
      cdst[cnt] = dd; for (int k = 0; k < 3; ++k) { cn[cnt][k] = n[k]; cp[cnt][k] = v[k] - n[k] * dd * real(0.5); }
      ++cnt;
    }
  } else if (kind == PAIR_SPHERE_BOX) {
    cnt = sphere_box(margin, g1, s1[0], g2, s2, cp[0], cn[0], cdst);
  } else if (kind == PAIR_CAPSULE_BOX) {   // closest point of the capsule axis to the box by alternating projections, then sphere vs box
    const real *ax = g1 + 3; const real hl = s1[1]; real t, p[3], pl[3], q[3], qw[3], tmp[3];
    for (int k = 0; k < 3; ++k) tmp[k] = g2[k] - g1[k];
    t = clampr(dot3(tmp, ax), -hl, hl);
    for (int it = 0; it < 4; ++it) {
      for (int k = 0; k < 3; ++k) { p[k] = g1[k] + ax[k] * t; tmp[k] = p[k] - g2[k]; }
      box_to_local(g2, tmp, pl); for (int k = 0; k < 3; ++k) q[k] = clampr(pl[k], -s2[k], s2[k]);
      box_to_world(g2, q, qw); for (int k = 0; k < 3; ++k) tmp[k] = g2[k] + qw[k] - g1[k];
      t = clampr(dot3(tmp, ax), -hl, hl);
    }
    for (int k = 0; k < 3; ++k) p[k] = g1[k] + ax[k] * t;
    cnt = sphere_box(margin, p, s1[0], g2, s2, cp[0], cn[0], cdst);
  } else {                                 // box - box: corners of g2 inside g1, then corners of g1 inside g2
    for (int pass = 0; pass < 2 && cnt < 4; ++pass) {
      const real *ga = pass ? g2 : g1, *gb = pass ? g1 : g2, *as = pass ? s2 : s1, *bs = pass ? s1 : s2;
      for (int i = 0; i < 8 && cnt < 4; ++i) {
        real v[3], vl[3]; box_corner(gb, bs, i, v);
        real t[3] = {v[0] - ga[0], v[1] - ga[1], v[2] - ga[2]}; box_to_local(ga, t, vl);
        int ax = -1; real best = real(1e30);
        for (int k = 0; k < 3; ++k) { const real pen = as[k] - mabs(vl[k]); if (pen <= -margin) { ax = -1; break; } if (pen < best) { best = pen; ax = k; } }
        if (ax < 0) continue;
        real nl[3] = {0, 0, 0}, nw[3]; nl[ax] = vl[ax] >= 0 ? real(1) : real(-1); box_to_world(ga, nl, nw);
        const real sgn = pass ? real(-1) : real(1);
        cdst[cnt] = -best; for (int k = 0; k < 3; ++k) { cn[cnt][k] = sgn * nw[k]; cp[cnt][k] = v[k] + nw[k] * best * real(0.5); }
        ++cnt;
      }
    }
  }
  return cnt;
}

// translational Jacobian column of dof l for a world point attached to `body` (zero when l is not in the body's chain)
template <typename real> CFN void jac_col(const DevModel<real> &cm, int l, int body, const real *cd, const real *point, const real *com, real *out) {
  if ((cm.body_dofmask[body] >> l) & 1u) {
    real off[3] = {point[0] - com[0], point[1] - com[1], point[2] - com[2]}, t[3];
    cross3(t, cd, off);
    out[0] = cd[3] + t[0]; out[1] = cd[4] + t[1]; out[2] = cd[5] + t[2];
  } else { out[0] = out[1] = out[2] = 0; }
}

// sliding friction of a candidate pair: the shared model's value, or mixed from the env's per-geom values (mj_contactParam: the geom of
// higher priority wins, else the larger coefficient)
template <typename real>
CFN real pair_friction(const DevModel<real> &cm, const real *ce, int p) {
  const uint32_t code = cm.pair_code[p]; const int pc = pair_pc(code);
  if (!ce) return cm.pc_mu[pc];
  const real f1 = ce[CE_FRIC + pair_g1(code)], f2 = ce[CE_FRIC + pair_g2(code)]; const int src = cm.pc_mu_src[pc];
  return src == 0 ? mmax(f1, f2) : (src == 1 ? f1 : f2);
}
// one row of J  ->  D^-1/2 L^-T J' in place (mj_solveM2 on a single vector, main-tree dofs); returns its squared norm = J inv(M) J'
template <typename real>
CFN real half_solve_row(const DevModel<real> &cm, const real *sm, real *yy, int nv) {
  const real *qLD = sm + S_QLD;
  for (int i = nv - 1; i > 0; --i) {
    const int di = cm.dof_depth[i]; if (di == 0) continue;
    const real xi = yy[i]; const real *Li = qLD + cm.dof_Madr[i]; const unsigned char *an = cm.dof_anc[i];
    for (int t = 1; t <= di; ++t) yy[an[t]] -= Li[t] * xi;
  }
  real ad = 0;
  for (int d = 0; d < nv; ++d) { const real v = yy[d] * sm[S_DSQI + d]; yy[d] = v; ad += v * v; }
  return ad;
}

// the same transform for a row that touches the base chain and ONE of the two mirrored legs (off = 0: first leg, off = sym_n: second leg): the
// leg's dofs are walked through the first leg's tables with a per-lane offset, the other leg's entries (exact zeros) are left alone
template <typename real>
CFN real half_solve_row_sym(const DevModel<real> &cm, const real *sm, real *yy, int off) {
  const real *qLD = sm + S_QLD; const int f = cm.sym_first, n = cm.sym_n, moff = off ? cm.sym_madr : 0;
  for (int r = n - 1; r >= 0; --r) {
    const int i0 = f + r, di = cm.dof_depth[i0];
    const real xi = yy[i0 + off]; const real *Li = qLD + cm.dof_Madr[i0] + moff; const unsigned char *an = cm.dof_anc[i0];
    for (int t = 1; t <= di; ++t) { int a = an[t]; if (a >= f) a += off; yy[a] -= Li[t] * xi; }
  }
  for (int i = f - 1; i > 0; --i) {
    const int di = cm.dof_depth[i]; const real xi = yy[i]; const real *Li = qLD + cm.dof_Madr[i]; const unsigned char *an = cm.dof_anc[i];
    for (int t = 1; t <= di; ++t) yy[an[t]] -= Li[t] * xi;
  }
  real ad = 0;
  for (int d = 0; d < f; ++d) { const real v = yy[d] * sm[S_DSQI + d]; yy[d] = v; ad += v * v; }
  for (int r = 0; r < n; ++r) { const int d = f + r + off; const real v = yy[d] * sm[S_DSQI + d]; yy[d] = v; ad += v * v; }
  return ad;
}
CFN int side_of(int s1, int s2) { return s1 == 0 ? s2 : (s2 == 0 || s2 == s1 ? s1 : 3); }

// cassie_sim_foot_velocities (src/cassiemujoco.c:1623-1631): mj_comVel of the two foot bodies = sum over the dof chain, root first, of
// cdof * qvel; qvel is read from vecs[0..nv), cdof from the copy the derived-quantity stage left in the geom buffer
template <typename real>
CFN void aux_foot_velocities(const DevModel<real> &cm, real *sm, real *aux) {
  DECL_LANE
  const real *vecs = sm + S_VEC, *cdofs = sm + scratch_reals(cm.ystride);   // the extended instance's tail
  LANES
    if (l < 12) {
      const int fb = cm.foot_body[l / 6], k = l % 6; real acc = 0;
      if (fb >= 0) {
        const int ld = cm.body_lastdof[fb];
        if (ld >= 0) { for (int t = cm.dof_depth[ld]; t >= 1; --t) { const int a = cm.dof_anc[ld][t]; acc += cdofs[6 * a + k] * vecs[a]; } acc += cdofs[6 * ld + k] * vecs[ld]; }
      }
      aux[AX_FOOT_VEL + l] = acc;
    }
  ENDL
}

// ------------------------------------------------------------------ one MuJoCo sub-step (mj_step1 + mj_step2)
// state in: sm[S_QPOS], lane vars qvel / qacc_ws, ctrl in sm[S_CST..] (via ctrl lane var), xfrc.  state out: same + sensordata.
// DR ("extended" instance): the batch carries per-environment model constants (domain randomisation) and / or derived-quantity rows.  A
// compile-time flag so that the plain instance keeps its constants in the shared model block with no indirection and contains neither the
// set_const stage nor the derived-quantity stages.
template <typename real, bool DR, int FEAT>
CFN void mj_substep(const DevModel<real> &cm, real *sm, const EnvPtrs<real> &E, LP(real, qvel), LP(real, qacc_ws), LP(real, xqvel), LP(real, xqacc_ws), LP(real, ctrl), real *dbg, real *aux_row, int mode) {
  const bool advance = (mode == 0);   // mode: 0 step, 1 mj_forward only, 2 query (kinematics + velocities -> centre-of-mass slots of the aux row, nothing else written),
                                      //       3 set_const (invweights / masses / mean inertia at the reference configuration -> the env's constant row)
  // model constants that may be env-private (domain randomisation): the env's row if the batch carries one, else the shared block
  const real *ce = DR ? E.cenv : (const real *)0;
  real *const auxr = DR ? aux_row : (real *)0;   // derived-quantity row: only the extended instance carries those stages, and only the last
                                                 // sub-step of a launch fills it (what a query after the launch would see)
  const real *bmass = ce ? ce + CE_MASS : cm.body_mass, *bipos = ce ? ce + CE_IPOS : &cm.body_ipos[0][0], *ddamp = ce ? ce + CE_DAMP : cm.dof_damping;
  const real *binvw = ce ? ce + CE_BINVW : cm.body_invw, *dinvw = ce ? ce + CE_DINVW : cm.dof_invweight0;
  const bool use2 = advance && (cm.has_damping || ce);   // the Euler stage needs the factor of M + h B: produced together with M's
  const bool keep_qM = dbg || mode == 3;                 // the unfactored M itself is only wanted by the debug dump and by set_const
  const real root_mass_inv = ce ? ce[CE_ROOT_MINV] : cm.root_mass_inv, total_mass_inv = ce ? ce[CE_TOT_MINV] : cm.total_mass_inv, pgs_scale = ce ? ce[CE_PGS_SCALE] : cm.pgs_scale;
  const real xb_dsqi_t = ((FEAT & F_XB) && ce && cm.xb >= 0) ? real(1) / msqrt(bmass[cm.xb]) : cm.xb_dsqi[0];   // the extra free body's mass acts at once, like every body_mass entry
  DECL_LANE
  const int csync = E.cta_sync;   // bit k: rendezvous k of the sub-step is on (set by the kernel wrapper for multi-tick step launches only)
  STAGE_SYNC(csync & 1);
  // nv: dofs of the main tree (one per lane); xb: extra free body or -1.  An instance compiled without the extra-body feature sees xb = -1 and the
  // narrow row stride as constants: every `xb >= 0` branch below and the stride arithmetic fold away
  const int nv = cm.nv, nb = cm.nbody, ys = (FEAT & F_XB) ? cm.ystride : YSTRIDE_MAIN, xb = (FEAT & F_XB) ? cm.xb : -1;
  real *xpos = sm + S_XPOS, *xquat = sm + S_XQUAT, *xmat = sm + S_XMAT, *cdof = sm + S_CDOF;
  real *qpos = sm + S_QPOS, *qM = E.qM, *qLD = sm + S_QLD, *Y = sm + S_Y, *efc = sm + S_EFC, *con = sm + S_CON;
  real *qLD2 = sm + S_Y + T_CVEL;   // second matrix of the fused factorisation: the velocity-stage temporaries are not live yet
  const real *xfrc = E.xfrc; int *counters = E.counters;
  // temporaries of the smooth-dynamics stages live in the (not yet used) constraint-matrix region
  real *cinert = sm + S_Y + T_CINERT, *cdofd = sm + S_Y + T_CDOFD;
  real *xanchor = sm + S_Y + T_CRB, *xaxis = sm + S_Y + T_CRB + 96, *qloc = sm + S_Y + T_CRB + 192;
  LV(real, com0); LV(real, com1); LV(real, com2);

  // ================= kinematics (mj_kinematics) =================
  // (A) every body in parallel: its transform relative to the parent frame from its own joints (slides first, then at most one hinge /
  //     ball -- checked when the model block is built); (B) tree levels: compose with the parent; anchors / axes follow in the cdof phase
  real *prel = xanchor, *qrel = qloc, *sax = xaxis;   // [32][3] position, [32][4] rotation, [32][3] slide axes in the parent frame
  LANES
    if (l >= 1 && l < nb && l != xb) {
      real p[3] = {cm.body_pos[l][0], cm.body_pos[l][1], cm.body_pos[l][2]}, q[4] = {cm.body_quat[l][0], cm.body_quat[l][1], cm.body_quat[l][2], cm.body_quat[l][3]};
      for (int jj = 0; jj < cm.body_jntnum[l]; ++jj) {
        const int j = cm.body_jntadr[l] + jj, t = cm.jnt_type[j], qa = cm.jnt_qposadr[j];
        if (t == 2) {          // slide along its axis in the (not yet rotated) body frame
          real R[9], ax[3]; quat2mat(R, q); mat_vec(ax, R, cm.jnt_axis[j]);
          sax[3 * j] = ax[0]; sax[3 * j + 1] = ax[1]; sax[3 * j + 2] = ax[2];
          const real d = qpos[qa] - cm.jnt_qpos0[j];
          p[0] += ax[0] * d; p[1] += ax[1] * d; p[2] += ax[2] * d;
        } else {               // hinge / ball about the joint anchor
          real ql[4];
          if (t == 3) { real sn, cs; msincos(real(0.5) * (qpos[qa] - cm.jnt_qpos0[j]), &sn, &cs); ql[0] = cs; ql[1] = cm.jnt_axis[j][0] * sn; ql[2] = cm.jnt_axis[j][1] * sn; ql[3] = cm.jnt_axis[j][2] * sn; }
          else { ql[0] = qpos[qa]; ql[1] = qpos[qa + 1]; ql[2] = qpos[qa + 2]; ql[3] = qpos[qa + 3]; normalize4(ql); }
          if (cm.any_jnt_pos) {  // off-centre joint: the anchor stays fixed while the body turns about it
            real R[9], v0[3], v1[3]; quat2mat(R, q); mat_vec(v0, R, cm.jnt_pos[j]); mul_quat(q, q, ql); quat2mat(R, q); mat_vec(v1, R, cm.jnt_pos[j]);
            p[0] += v0[0] - v1[0]; p[1] += v0[1] - v1[1]; p[2] += v0[2] - v1[2];
          } else mul_quat(q, q, ql);
        }
      }
      prel[3 * l] = p[0]; prel[3 * l + 1] = p[1]; prel[3 * l + 2] = p[2];
      qrel[4 * l] = q[0]; qrel[4 * l + 1] = q[1]; qrel[4 * l + 2] = q[2]; qrel[4 * l + 3] = q[3];
    }
    if (l == 0) { xpos[0] = xpos[1] = xpos[2] = 0; xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0; for (int k = 0; k < 9; ++k) xmat[k] = (k % 4 == 0) ? real(1) : real(0); }
  ENDL
  for (int lev = 1; lev <= cm.maxdepth; ++lev) {
    LANES  // lane = body at this tree level
      if (l < nb && cm.body_depth[l] == lev) {
        real pos[3], quat[4];
        if (l == xb) {   // free joint: pose straight from qpos (position, quaternion)
          const int qa = cm.xb_qadr;
          pos[0] = qpos[qa]; pos[1] = qpos[qa + 1]; pos[2] = qpos[qa + 2];
          quat[0] = qpos[qa + 3]; quat[1] = qpos[qa + 4]; quat[2] = qpos[qa + 5]; quat[3] = qpos[qa + 6];
        } else {
          const int p = cm.body_parent[l]; real v[3];
          mat_vec(v, xmat + 9 * p, prel + 3 * l);
          pos[0] = xpos[3 * p] + v[0]; pos[1] = xpos[3 * p + 1] + v[1]; pos[2] = xpos[3 * p + 2] + v[2];
          mul_quat(quat, xquat + 4 * p, qrel + 4 * l);
        }
        normalize4(quat);
        xpos[3 * l] = pos[0]; xpos[3 * l + 1] = pos[1]; xpos[3 * l + 2] = pos[2];
        xquat[4 * l] = quat[0]; xquat[4 * l + 1] = quat[1]; xquat[4 * l + 2] = quat[2]; xquat[4 * l + 3] = quat[3];
        quat2mat(xmat + 9 * l, quat);
      }
    ENDL
  }

  // ================= comPos: subtree com of the root, cinert, cdof =================
  LV(real, t0); LV(real, t1); LV(real, t2);
  LANES  // lane = body
    L(t0) = L(t1) = L(t2) = 0;
    if (l >= 1 && l < nb && l != xb) {
      real v[3]; mat_vec(v, xmat + 9 * l, bipos + 3 * l);
      const real m = bmass[l];
      L(t0) = m * (xpos[3 * l] + v[0]); L(t1) = m * (xpos[3 * l + 1] + v[1]); L(t2) = m * (xpos[3 * l + 2] + v[2]);
    }
  ENDL
  ALLSUM(t0); ALLSUM(t1); ALLSUM(t2);
  LANES L(com0) = L(t0) * root_mass_inv; L(com1) = L(t1) * root_mass_inv; L(com2) = L(t2) * root_mass_inv; ENDL
  LANES  // lane = body: cinert about the com, world orientation
    if (l < nb) {
      real *r = cinert + 10 * l;
      if (l == 0) { for (int k = 0; k < 10; ++k) r[k] = 0; }
      else {
        real v[3], off[3], R[9]; const real *X = xmat + 9 * l, *B = cm.body_imat[l], *I = cm.body_inertia[l]; const real m = bmass[l];
        mat_vec(v, X, bipos + 3 * l);
        off[0] = xpos[3 * l] + v[0] - L(com0); off[1] = xpos[3 * l + 1] + v[1] - L(com1); off[2] = xpos[3 * l + 2] + v[2] - L(com2);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) R[3 * a + b] = X[3 * a] * B[b] + X[3 * a + 1] * B[3 + b] + X[3 * a + 2] * B[6 + b];
        r[0] = R[0] * R[0] * I[0] + R[1] * R[1] * I[1] + R[2] * R[2] * I[2] + m * (off[1] * off[1] + off[2] * off[2]);
        r[1] = R[3] * R[3] * I[0] + R[4] * R[4] * I[1] + R[5] * R[5] * I[2] + m * (off[0] * off[0] + off[2] * off[2]);
        r[2] = R[6] * R[6] * I[0] + R[7] * R[7] * I[1] + R[8] * R[8] * I[2] + m * (off[0] * off[0] + off[1] * off[1]);
        r[3] = R[0] * R[3] * I[0] + R[1] * R[4] * I[1] + R[2] * R[5] * I[2] - m * off[0] * off[1];
        r[4] = R[0] * R[6] * I[0] + R[1] * R[7] * I[1] + R[2] * R[8] * I[2] - m * off[0] * off[2];
        r[5] = R[3] * R[6] * I[0] + R[4] * R[7] * I[1] + R[5] * R[8] * I[2] - m * off[1] * off[2];
        r[6] = m * off[0]; r[7] = m * off[1]; r[8] = m * off[2]; r[9] = m;
      }
    }
  ENDL
  LV(real, cd0); LV(real, cd1); LV(real, cd2); LV(real, cd3); LV(real, cd4); LV(real, cd5);  // this lane's cdof (angular; linear)
  LANES  // lane = dof
    L(cd0) = L(cd1) = L(cd2) = L(cd3) = L(cd4) = L(cd5) = 0;
    if (l < nv) {
      const int j = cm.dof_jnt[l], t = cm.jnt_type[j], b = cm.dof_body[l];
      real ax[3], c[3];
      if (t == 2) {  // slide: axis fixed in the parent frame
        mat_vec(ax, xmat + 9 * cm.body_parent[b], sax + 3 * j);
        L(cd3) = ax[0]; L(cd4) = ax[1]; L(cd5) = ax[2];
      } else {
        real an[3] = {xpos[3 * b], xpos[3 * b + 1], xpos[3 * b + 2]};
        if (cm.any_jnt_pos) { real v[3]; mat_vec(v, xmat + 9 * b, cm.jnt_pos[j]); an[0] += v[0]; an[1] += v[1]; an[2] += v[2]; }
        real off[3] = {L(com0) - an[0], L(com1) - an[1], L(com2) - an[2]};
        if (t == 3) mat_vec(ax, xmat + 9 * b, cm.jnt_axis[j]);
        else { const int k = l - cm.jnt_dofadr[j]; ax[0] = xmat[9 * b + k]; ax[1] = xmat[9 * b + 3 + k]; ax[2] = xmat[9 * b + 6 + k]; }
        cross3(c, ax, off);
        L(cd0) = ax[0]; L(cd1) = ax[1]; L(cd2) = ax[2]; L(cd3) = c[0]; L(cd4) = c[1]; L(cd5) = c[2];
      }
      real *o = cdof + 6 * l; o[0] = L(cd0); o[1] = L(cd1); o[2] = L(cd2); o[3] = L(cd3); o[4] = L(cd4); o[5] = L(cd5);
    }
  ENDL

  // ================= CRB -> sparse qM (mj_crb); crb overwrites the xanchor/xaxis/qloc temporaries =================
  {
    real *crb = sm + S_Y + T_CRB;
    LANES  // lane = body: composite inertia = sum of cinert over the (contiguous, depth-first) subtree
      if (l < nb) {
        real acc[10]; for (int k = 0; k < 10; ++k) acc[k] = 0;
        if (l >= 1) { const int end = cm.body_kid_dofs[l] ? l + 1 : cm.body_subtree_end[l]; for (int b = l; b < end; ++b) for (int k = 0; k < 10; ++k) acc[k] += cinert[10 * b + k]; }
        for (int k = 0; k < 10; ++k) crb[10 * l + k] = acc[k];
      }
    ENDL
    if (cm.any_big) {   // a body with a large subtree (the pelvis): its own term + the finished sums of its direct children (mj_crb's own order of accumulation)
      LANES
        if (l < nb && cm.body_kid_dofs[l]) {
          for (uint32_t kd = cm.body_kid_dofs[l]; kd; kd >>= 8) { const int c = cm.dof_body[(kd & 255u) - 1]; for (int k = 0; k < 10; ++k) crb[10 * l + k] += crb[10 * c + k]; }
        }
      ENDL
    }
    LANES  // lane = dof i: M(i, ancestors)
      if (l < nv) {
        real buf[6], cd[6] = {L(cd0), L(cd1), L(cd2), L(cd3), L(cd4), L(cd5)};
        mul_inert_vec(buf, crb + 10 * cm.dof_body[l], cd);
        int a = cm.dof_Madr[l];
        for (int j = l; j >= 0; j = cm.dof_parent[j]) {
          const real *cj = cdof + 6 * j;
          real s = cj[0] * buf[0] + cj[1] * buf[1] + cj[2] * buf[2] + cj[3] * buf[3] + cj[4] * buf[4] + cj[5] * buf[5];
          if (j == l) s += cm.dof_armature[l];
          qLD[a] = s; if (keep_qM) qM[a] = s;
          if (use2) qLD2[a] = (j == l) ? s + cm.timestep * ddamp[l] : s;
          ++a;
        }
      }
    ENDL
  }
  factor_ld(cm, sm, use2 ? qLD2 : (real *)0);
  // the factor of M + h B waits in the env's global scratch row until the Euler stage (the constraint stage needs this shared memory)
  if (use2) { LANES for (int k = l; k < cm.nM; k += 32) qM[NM_MAX + k] = qLD2[k]; ENDL }
  if (dbg) { LANES for (int a = l; a < cm.nM; a += 32) { dbg[D_QM + a] = qM[a]; dbg[D_QLD + a] = qLD[a]; } ENDL }

  // ================= mode 3: mj_setConst for this env (src/cassiemujoco.c:949-977 -> set0 / setStat [M]) =================
  // the caller loaded qpos0; with M factored there: dof_invweight0 = diag(inv(M)) (averaged over a ball joint's three dofs),
  // body_invweight0 = trace(Jp inv(M) Jp')/3 with Jp the translational Jacobian at the body's com, each entry the squared norm of a
  // half-solved row (the same in-place transform the constraint stage applies to J); mean inertia = mean diagonal of M
  if (DR && mode == 3) {
    real *cew = E.cenv, *vecs3 = sm + S_VEC;
    if (!cew) return;
    LANES for (int d = 0; d < ys; ++d) Y[l * ys + d] = (d == l) ? real(1) : real(0); ENDL
    LANES if (l < nv) vecs3[l] = half_solve_row(cm, sm, Y + l * ys, nv); ENDL
    LANES
      if (l < nv) {
        const int j = cm.dof_jnt[l], da = cm.jnt_dofadr[j];
        cew[CE_DINVW + l] = (cm.jnt_type[j] == 1) ? (vecs3[da] + vecs3[da + 1] + vecs3[da + 2]) / 3 : vecs3[l];
      }
    ENDL
    for (int c0 = 1; c0 < nb; c0 += 16) {
      const int nbc = (nb - c0) < 16 ? (nb - c0) : 16;
      for (int i = 0; i < nbc; ++i) {
        const int b = c0 + i;
        LANES  // lane = dof: column of the translational Jacobian at the body's centre of mass
          real v[3], pt[3], jc[3] = {0, 0, 0}, cd[6] = {L(cd0), L(cd1), L(cd2), L(cd3), L(cd4), L(cd5)}, cm3[3] = {L(com0), L(com1), L(com2)};
          if (b != xb) { mat_vec(v, xmat + 9 * b, bipos + 3 * b); pt[0] = xpos[3 * b] + v[0]; pt[1] = xpos[3 * b + 1] + v[1]; pt[2] = xpos[3 * b + 2] + v[2]; jac_col(cm, l, b, cd, pt, cm3, jc); }
          for (int k = 0; k < 3; ++k) Y[(3 * i + k) * ys + l] = jc[k];
        ENDL
      }
      LANES  // lane = row
        if (l < 3 * nbc) vecs3[l] = half_solve_row(cm, sm, Y + l * ys, nv);
        if (l + 32 < 3 * nbc) vecs3[l + 32] = half_solve_row(cm, sm, Y + (l + 32) * ys, nv);
      ENDL
      LANES
        if (l < nbc) { const int b = c0 + l; cew[CE_BINVW + b] = (b == xb) ? real(1) / bmass[b] : (cm.body_lastdof[b] >= 0 ? (vecs3[3 * l] + vecs3[3 * l + 1] + vecs3[3 * l + 2]) / 3 : real(0)); }
      ENDL
    }
    LV(real, sm0); LV(real, sm1);
    LANES L(sm0) = (l >= 1 && l < nb && l != xb) ? bmass[l] : real(0); L(sm1) = (l < nv) ? qM[cm.dof_Madr[l]] : real(0); ENDL
    ALLSUM(sm0); ALLSUM(sm1);
    LANES
      if (l == 0) {
        real mroot = L(sm0), mtot = mroot, tr = L(sm1); int nvt = nv;
        if (xb >= 0) { mtot += bmass[xb]; tr += 3 * bmass[xb] + cm.xb_inertia[0] + cm.xb_inertia[1] + cm.xb_inertia[2]; nvt += 6; }
        cew[CE_ROOT_MINV] = real(1) / mroot; cew[CE_TOT_MINV] = real(1) / mtot;
        cew[CE_PGS_SCALE] = real(1) / ((tr / (real)nvt) * (real)nvt);
      }
    ENDL
    return;
  }

  // ================= velocity stage: comVel, passive, RNE bias =================
  real *S = sm + S_Y + T_CRB;      // chain sums [32][6] (crb is dead)
  real *cvel = sm + S_Y + T_CVEL;  // [32][6]
  real *cfrc = sm + S_Y + T_CFRC;  // [32][6]
  real *vecs = sm + S_VEC;
  LANES if (l < nv) vecs[l] = L(qvel); ENDL
  // chain sums over a dof's ancestors: with two legs below a common base chain (sym_on) a leg dof walks its own leg only and then takes the base chain's
  // finished sum (the base's last dof), which is also the root-first order mj_comVel accumulates in
#ifdef CASSIE_NO_CHAIN_CUT
  const int chain_cut = -1;
#else
  const int chain_cut = cm.sym_on ? cm.sym_first - 1 : -1;
#endif
  LANES  // lane = dof: S_d = sum over the chain root..d of cdof_a * qvel_a
    if (l < nv) {
      real acc[6] = {0, 0, 0, 0, 0, 0}; const int stop = l > chain_cut ? chain_cut : -1;
      for (int a = l; a > stop; a = cm.dof_parent[a]) { const real q = vecs[a]; const real *c = cdof + 6 * a; for (int k = 0; k < 6; ++k) acc[k] += c[k] * q; }
      for (int k = 0; k < 6; ++k) S[6 * l + k] = acc[k];
    }
  ENDL
  if (chain_cut >= 0) { LANES if (l < nv && l > chain_cut) for (int k = 0; k < 6; ++k) S[6 * l + k] += S[6 * chain_cut + k]; ENDL }
  LANES  // cdof_dot = (velocity before this joint) x cdof ; body velocities
    if (l < nv) {
      real vb[6] = {0, 0, 0, 0, 0, 0}, cd[6] = {L(cd0), L(cd1), L(cd2), L(cd3), L(cd4), L(cd5)}, r[6];
      const int src = cm.dof_cvelsrc[l];
      if (src >= 0) for (int k = 0; k < 6; ++k) vb[k] = S[6 * src + k];
      cross_motion(r, vb, cd);
      for (int k = 0; k < 6; ++k) cdofd[6 * l + k] = r[k];
    }
    if (l < nb) { const int ld = cm.body_lastdof[l]; for (int k = 0; k < 6; ++k) cvel[6 * l + k] = (ld >= 0) ? S[6 * ld + k] : real(0); }
  ENDL
  LANES  // T_d = chain sums of cdof_dot * qvel (overwrites S; nobody reads S in this phase)
    if (l < nv) {
      real acc[6] = {0, 0, 0, 0, 0, 0}; const int stop = l > chain_cut ? chain_cut : -1;
      for (int a = l; a > stop; a = cm.dof_parent[a]) { const real q = vecs[a]; const real *c = cdofd + 6 * a; for (int k = 0; k < 6; ++k) acc[k] += c[k] * q; }
      for (int k = 0; k < 6; ++k) S[6 * l + k] = acc[k];
    }
  ENDL
  if (chain_cut >= 0) { LANES if (l < nv && l > chain_cut) for (int k = 0; k < 6; ++k) S[6 * l + k] += S[6 * chain_cut + k]; ENDL }
  LANES  // lane = body: cfrc = cinert * cacc + cvel x* (cinert * cvel), cacc = (0; -g) + T[lastdof]
    if (l < nb) {
      real f[6] = {0, 0, 0, 0, 0, 0};
      if (l >= 1) {
        real acc[6] = {0, 0, 0, -cm.gravity[0], -cm.gravity[1], -cm.gravity[2]}, tmp[6], tmp1[6];
        const int ld = cm.body_lastdof[l];
        if (ld >= 0) for (int k = 0; k < 6; ++k) acc[k] += S[6 * ld + k];
        mul_inert_vec(f, cinert + 10 * l, acc);
        mul_inert_vec(tmp, cinert + 10 * l, cvel + 6 * l);
        cross_force(tmp1, cvel + 6 * l, tmp);
        for (int k = 0; k < 6; ++k) f[k] += tmp1[k];
      }
      for (int k = 0; k < 6; ++k) cfrc[6 * l + k] = f[k];
    }
  ENDL
  LANES { const int ld = cm.body_lastdof[cm.imu_body]; if (l < 6) vecs[96 + l] = (ld >= 0) ? S[6 * ld + l] : real(0); } ENDL
  LV(real, qfrc_smooth); LV(real, qacc_smooth); LV(real, bias);
  // lane = dof: the sum of cfrc over the subtree of the dof's body (a body with a large subtree: its own term, then the sums its children's first dof lanes hold)
  LV(real, sf0); LV(real, sf1); LV(real, sf2); LV(real, sf3); LV(real, sf4); LV(real, sf5);
  LANES
    real acc[6] = {0, 0, 0, 0, 0, 0};
    if (l < nv) { const int b = cm.dof_body[l], end = cm.body_kid_dofs[b] ? b + 1 : cm.body_subtree_end[b]; for (int d = b; d < end; ++d) for (int k = 0; k < 6; ++k) acc[k] += cfrc[6 * d + k]; }
    L(sf0) = acc[0]; L(sf1) = acc[1]; L(sf2) = acc[2]; L(sf3) = acc[3]; L(sf4) = acc[4]; L(sf5) = acc[5];
  ENDL
  if (cm.any_big) {
    LV(uint32_t, kd); LV(int, ksrc); LV(int, more); LV(real, g);
    LANES_NS L(kd) = l < nv ? cm.body_kid_dofs[cm.dof_body[l]] : 0u; ENDL_NS
    for (int j = 0; j < 4; ++j) {
      LANES_NS L(ksrc) = (L(kd) & 255u) ? (int)(L(kd) & 255u) - 1 : l; ENDL_NS
#define KID_ADD_(v) SHFLV(g, v, L(ksrc)); LANES_NS if (L(kd) & 255u) L(v) += L(g); ENDL_NS
      KID_ADD_(sf0) KID_ADD_(sf1) KID_ADD_(sf2) KID_ADD_(sf3) KID_ADD_(sf4) KID_ADD_(sf5)
#undef KID_ADD_
      LANES_NS L(kd) >>= 8; L(more) = L(kd) != 0u; ENDL_NS
      uint32_t anymore; BALLOT(anymore, more); if (!anymore) break;   // the Cassie pelvis has two such children
    }
  }
  LANES  // lane = dof: bias = cdof . (that sum); passive; actuation; applied
    L(qfrc_smooth) = 0; L(bias) = 0;
    if (l < nv) {
      const int b = cm.dof_body[l];
      const real bs = L(cd0) * L(sf0) + L(cd1) * L(sf1) + L(cd2) * L(sf2) + L(cd3) * L(sf3) + L(cd4) * L(sf4) + L(cd5) * L(sf5);
      const int j = cm.dof_jnt[l];
      real passive = -ddamp[l] * L(qvel);
      if (cm.jnt_stiffness[j] != 0 && cm.jnt_type[j] >= 2) passive -= cm.jnt_stiffness[j] * (qpos[cm.jnt_qposadr[j]] - cm.jnt_qspring[j]);
      real f = passive - bs;
      // xfrc_applied on one body: J(xipos)' [force; torque]
      const int xb = (int)xfrc[6];
      if (xb > 0 && ((cm.body_dofmask[xb] >> l) & 1u)) {
        real v[3], pt[3], jc[3], cd[6] = {L(cd0), L(cd1), L(cd2), L(cd3), L(cd4), L(cd5)}, cm3[3] = {L(com0), L(com1), L(com2)};
        mat_vec(v, xmat + 9 * xb, bipos + 3 * xb);
        pt[0] = xpos[3 * xb] + v[0]; pt[1] = xpos[3 * xb + 1] + v[1]; pt[2] = xpos[3 * xb + 2] + v[2];
        jac_col(cm, l, xb, cd, pt, cm3, jc);
        f += jc[0] * xfrc[0] + jc[1] * xfrc[1] + jc[2] * xfrc[2] + cd[0] * xfrc[3] + cd[1] * xfrc[4] + cd[2] * xfrc[5];
      }
      L(bias) = bs;
      L(qfrc_smooth) = f;
      if (dbg) { dbg[D_BIAS + l] = bs; dbg[D_PASSIVE + l] = passive; }
    }
  ENDL
  LANES  // motors: lane = actuator; clamp ctrl, scatter gear * ctrl to the dof through shared memory
    vecs[32 + l] = 0;
  ENDL
  LANES if (l < cm.nu) vecs[32 + cm.act_dof[l]] = cm.act_gear[l] * clampr(L(ctrl), cm.act_ctrl_lo[l], cm.act_ctrl_hi[l]); ENDL
  LANES if (l < nv) { L(qfrc_smooth) += vecs[32 + l]; L(qacc_smooth) = L(qfrc_smooth); } else L(qacc_smooth) = 0; ENDL
  solve_m(cm, sm, qacc_smooth);
  // ---- extra free body (cassie_tray_box.xml: the cup): its mass matrix is the constant diag(m, m, m, I1, I2, I3) (translation in the
  // world frame, rotation in the body frame, inertial frame = body frame), so its smooth acceleration is closed form; lanes 0..5 = its dofs
  LV(real, xqacc_smooth); LV(real, xz); LV(real, xqacc);
  LANES L(xqacc_smooth) = 0; L(xz) = 0; L(xqacc) = 0; if (xb >= 0 && l < 6) vecs[160 + l] = L(xqvel); ENDL
  if (xb >= 0) {
    LANES
      if (l < 3) L(xqacc_smooth) = cm.gravity[l];
      else if (l < 6) {
        const real w[3] = {vecs[163], vecs[164], vecs[165]}, Iw[3] = {cm.xb_inertia[0] * w[0], cm.xb_inertia[1] * w[1], cm.xb_inertia[2] * w[2]}; real t3[3];
        cross3(t3, w, Iw);
        L(xqacc_smooth) = -t3[l - 3] / cm.xb_inertia[l - 3];
      }
    ENDL
  }
  if (dbg) {
    LANES
      if (l < nv) { dbg[D_SMOOTH + l] = L(qfrc_smooth); dbg[D_QACCS + l] = L(qacc_smooth); for (int k = 0; k < 6; ++k) dbg[D_CDOF + 6 * l + k] = cdof[6 * l + k]; }
      if (l < nb) { for (int k = 0; k < 3; ++k) dbg[D_XPOS + 3 * l + k] = xpos[3 * l + k]; for (int k = 0; k < 4; ++k) dbg[D_XQUAT + 4 * l + k] = xquat[4 * l + k]; }
    ENDL
  }

  // ================= derived quantities, part 1: centre of mass, its velocity, angular momentum about it =================
  // (cassie_sim_cm_position / cm_velocity / angular_momentum, src/cassiemujoco.c:1633-1646,1693-1699 -> subtree_com, mj_subtreeVel of the
  // world body).  Sum over bodies of cinert * cvel is the spatial momentum about the main tree's com; the extra free body is folded in.
  if (auxr) {
    LV(real, m0); LV(real, m1); LV(real, m2); LV(real, m3); LV(real, m4); LV(real, m5);
    LANES
      real h6[6] = {0, 0, 0, 0, 0, 0};
      if (l >= 1 && l < nb && l != xb) mul_inert_vec(h6, cinert + 10 * l, cvel + 6 * l);
      L(m0) = h6[0]; L(m1) = h6[1]; L(m2) = h6[2]; L(m3) = h6[3]; L(m4) = h6[4]; L(m5) = h6[5];
    ENDL
    ALLSUM(m0); ALLSUM(m1); ALLSUM(m2); ALLSUM(m3); ALLSUM(m4); ALLSUM(m5);
    LANES
      if (l == 0) {
        real C[3] = {L(com0), L(com1), L(com2)}, P[3] = {L(m3), L(m4), L(m5)}, Lc[3] = {L(m0), L(m1), L(m2)}, V[3];
        if (xb >= 0) {
          const real mc = bmass[xb], Mr = real(1) / root_mass_inv, *xc = xpos + 3 * xb, *R = xmat + 9 * xb;
          real Ct[3], d1[3], d2[3], t3[3], pc[3] = {mc * vecs[160], mc * vecs[161], mc * vecs[162]}, Iw[3] = {cm.xb_inertia[0] * vecs[163], cm.xb_inertia[1] * vecs[164], cm.xb_inertia[2] * vecs[165]}, Lx[3];
          for (int k = 0; k < 3; ++k) { Ct[k] = (Mr * C[k] + mc * xc[k]) * total_mass_inv; d1[k] = C[k] - Ct[k]; d2[k] = xc[k] - Ct[k]; }
          mat_vec(Lx, R, Iw);
          cross3(t3, d1, P); for (int k = 0; k < 3; ++k) Lc[k] += t3[k] + Lx[k];
          cross3(t3, d2, pc); for (int k = 0; k < 3; ++k) { Lc[k] += t3[k]; P[k] += pc[k]; C[k] = Ct[k]; }
        }
        for (int k = 0; k < 3; ++k) V[k] = P[k] * total_mass_inv;
        for (int k = 0; k < 3; ++k) { auxr[AX_CM_POS + k] = C[k]; auxr[AX_CM_VEL + k] = V[k]; auxr[AX_ANGMOM + k] = Lc[k]; }
      }
    ENDL
    if (mode == 2) return;
  }

  // ================= sensors that do not need qacc (mj_sensorPos / mj_sensorVel for the Cassie layout) =================
  // done here because the constraint stage below reuses the kinematics buffers; the accelerometer is finished after the solve from
  // a small stash: vecs[102..104] = qacc-independent part, vecs[105..] = 3 x nchain gain matrix over the IMU body's dof chain
  {
    real *cst = E.cst;
    LANES
      if (l < 16) cst[CS_SENSOR + l] = cm.enc_scale[l] * qpos[cm.enc_qposadr[l]];   // actuatorpos = gear * q, jointpos = q
      if (l < cm.nu) cst[CS_ACTVEL + l] = cm.act_gear[l] * vecs[cm.act_dof[l]];
      if (l == 16) {  // IMU: framequat, gyro, magnetometer on site `imu`
        const int b = cm.imu_body; real q[4], Rs[9], sp[3], v[3];
        mul_quat(q, xquat + 4 * b, cm.imu_quat);
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) Rs[3 * a + c] = xmat[9 * b + 3 * a] * cm.imu_mat[c] + xmat[9 * b + 3 * a + 1] * cm.imu_mat[3 + c] + xmat[9 * b + 3 * a + 2] * cm.imu_mat[6 + c];
        mat_vec(v, xmat + 9 * b, cm.imu_pos); sp[0] = xpos[3 * b] + v[0]; sp[1] = xpos[3 * b + 1] + v[1]; sp[2] = xpos[3 * b + 2] + v[2];
        real dif[3] = {sp[0] - L(com0), sp[1] - L(com1), sp[2] - L(com2)};
        // body spatial velocity in the c-frame and the qacc-independent part of its acceleration (chain sums over the body's dofs)
        real cv[6] = {0, 0, 0, 0, 0, 0}, ca[6] = {0, 0, 0, -cm.gravity[0], -cm.gravity[1], -cm.gravity[2]};
        for (int k = 0; k < 6; ++k) ca[k] += vecs[96 + k];
        int nchain = 0;
        for (int a = cm.body_lastdof[b]; a >= 0; a = cm.dof_parent[a], ++nchain) {
          const real *cd_ = cdof + 6 * a; real t3[3], gl[3], gloc[3];
          for (int k = 0; k < 6; ++k) cv[k] += cd_[k] * vecs[a];
          cross3(t3, dif, cd_); gl[0] = cd_[3] - t3[0]; gl[1] = cd_[4] - t3[1]; gl[2] = cd_[5] - t3[2];
          matT_vec(gloc, Rs, gl);
          vecs[105 + 3 * nchain] = gloc[0]; vecs[105 + 3 * nchain + 1] = gloc[1]; vecs[105 + 3 * nchain + 2] = gloc[2];
        }
        vecs[127] = (real)nchain;
        real t3[3], vl[3], al[3], lw[3], lv[3], la[3], corr[3];
        cross3(t3, dif, cv); vl[0] = cv[3] - t3[0]; vl[1] = cv[4] - t3[1]; vl[2] = cv[5] - t3[2];
        cross3(t3, dif, ca); al[0] = ca[3] - t3[0]; al[1] = ca[4] - t3[1]; al[2] = ca[5] - t3[2];
        matT_vec(lw, Rs, cv); matT_vec(lv, Rs, vl); matT_vec(la, Rs, al); cross3(corr, lw, lv);
        for (int k = 0; k < 4; ++k) cst[CS_SENSOR + 16 + k] = q[k];
        for (int k = 0; k < 3; ++k) {
          real g = lw[k];
          if (cm.gyro_cutoff > 0) g = clampr(g, -cm.gyro_cutoff, cm.gyro_cutoff);
          cst[CS_SENSOR + 20 + k] = g; vecs[102 + k] = la[k] + corr[k];
        }
        real mg[3]; matT_vec(mg, Rs, cm.magnetic);
        for (int k = 0; k < 3; ++k) cst[CS_SENSOR + 26 + k] = mg[k];
      }
    ENDL
  }

  STAGE_SYNC(csync & 2);
  // ================= collision (lane = candidate geom pair) =================
  real *geom = sm + S_Y + T_GEOM;   // the smooth-dynamics temporaries below it are dead; the constraint rows are written after the contact list
  LANES   // world poses of the geoms on moving bodies (static geoms carry theirs in the model block)
    if (l < cm.ngeom) {
      const int b = cm.geom_body[l]; real v[3];
      mat_vec(v, xmat + 9 * b, cm.geom_pos[l]);
      geom[12 * l] = xpos[3 * b] + v[0]; geom[12 * l + 1] = xpos[3 * b + 1] + v[1]; geom[12 * l + 2] = xpos[3 * b + 2] + v[2];
      // world rotation of the geom, stored by columns: z axis, then x and y (capsules / planes only need z; boxes need all three)
      for (int c = 0; c < 3; ++c) { real col[3] = {cm.geom_mat[l][c], cm.geom_mat[l][3 + c], cm.geom_mat[l][6 + c]}; mat_vec(v, xmat + 9 * b, col);
        const int o = c == 2 ? 3 : (c == 0 ? 6 : 9); geom[12 * l + o] = v[0]; geom[12 * l + o + 1] = v[1]; geom[12 * l + o + 2] = v[2]; }
    }
  ENDL
  // broad phase for static obstacle boxes (the 15 stair boxes of model/cassie.xml:232-246, parked 20 m away unless a caller places them): a box
  // farther from the robot's root body than the robot can reach cannot touch it, so its 9 candidate pairs are skipped without a narrow phase
  uint32_t box_near = 0xffffffffu;
  if ((FEAT & F_BOX) && cm.npair > cm.npair_a) {
    LV(int, nearb);
    LANES
      L(nearb) = 1;
      if (l < cm.ngeom_static && cm.geom_type[MG + l] == 6) {
        const real *w = cm.geom_wpose[l], *rp = xpos + 3 * cm.root_body; const real dx = w[0] - rp[0], dy = w[1] - rp[1], dz = w[2] - rp[2], rr = cm.robot_reach + cm.geom_rbound[MG + l];
        L(nearb) = (dx * dx + dy * dy + dz * dz <= rr * rr) ? 1 : 0;
      }
    ENDL
    BALLOT(box_near, nearb);
  }
  int ncon_total = 0, ncon_a = 0;
  for (int seg = 0; seg < 2; ++seg) {   // run 0: the pairs without a static box; run 1: the static-box pairs, only when a box is within reach
  const int seg0 = seg ? cm.npair_a : 0, seg1 = seg ? cm.npair : cm.npair_a;
  if (seg == 1) { ncon_a = ncon_total; if (!(FEAT & F_BOX) || seg0 == seg1 || !(box_near & (uint32_t)cm.static_box_mask)) break; }
  for (int pbase = seg0; pbase < seg1; pbase += 32) {   // 32 candidate pairs per pass, one per lane
  LV(int, ccount); LV(int, coff);
  LVA(real, cb, 28);   // up to 4 contacts of this lane's pair: [pos3 normal3 dist] each
  LV(real, ch0); LV(real, ch1); LV(real, ch2);
  LANES
    L(ccount) = 0; L(ch0) = L(ch1) = L(ch2) = 0;
    const int pidx = pbase + l;
    bool go = pidx < seg1;
    const uint32_t code = go ? cm.pair_code[pidx] : 0u;
    const int g1 = pair_g1(code), g2 = pair_g2(code), kind = pair_kind(code);
    if ((FEAT & F_BOX) && go && kind >= PAIR_SPHERE_BOX) {   // a static box out of the robot's reach (pairs with the extra free body are always tested)
      if (g1 >= MG && !((box_near >> (g1 - MG)) & 1u) && cm.geom_body[g2] != xb) go = false;
      if (g2 >= MG && !((box_near >> (g2 - MG)) & 1u) && cm.geom_body[g1] != xb) go = false;
    }
    if (go) {
      const real margin = cm.pc_margin[pair_pc(code)];
      const real *p1 = g1 < MG ? geom + 12 * g1 : cm.geom_wpose[g1 - MG], *a1 = p1 + 3, *p2 = g2 < MG ? geom + 12 * g2 : cm.geom_wpose[g2 - MG], *a2 = p2 + 3;
      real cp[4][3], cn[4][3], cdst[4]; int n = 0;
      if (kind == PAIR_PLANE_SPHERE || kind == PAIR_PLANE_CAPSULE) {
        const real r = cm.geom_size[g2][0], hl = (kind == PAIR_PLANE_CAPSULE) ? cm.geom_size[g2][1] : real(0);
        const int ne = (kind == PAIR_PLANE_CAPSULE) ? 2 : 1;
        for (int e = 0; e < ne; ++e) {
          const real sgn = e ? real(-1) : real(1);
          real sp[3] = {p2[0] + sgn * hl * a2[0], p2[1] + sgn * hl * a2[1], p2[2] + sgn * hl * a2[2]};
          real dd[3] = {sp[0] - p1[0], sp[1] - p1[1], sp[2] - p1[2]};
          const real cdist = dot3(dd, a1);
          if (cdist <= margin + r && cdist - r < margin) {
            cdst[n] = cdist - r; cn[n][0] = a1[0]; cn[n][1] = a1[1]; cn[n][2] = a1[2];
            const real s = cdst[n] * real(0.5) + r;
            cp[n][0] = sp[0] - a1[0] * s; cp[n][1] = sp[1] - a1[1] * s; cp[n][2] = sp[2] - a1[2] * s; ++n;
          }
        }
        if (kind == PAIR_PLANE_CAPSULE) { L(ch0) = a2[0]; L(ch1) = a2[1]; L(ch2) = a2[2]; }
      } else if ((FEAT & F_HFIELD) && (kind == PAIR_HFIELD_SPHERE || kind == PAIR_HFIELD_CAPSULE)) {
        const real r = cm.geom_size[g2][0], hl = (kind == PAIR_HFIELD_CAPSULE) ? cm.geom_size[g2][1] : real(0);
        const int ne = (kind == PAIR_HFIELD_CAPSULE) ? 2 : 1;
        if (E.hfield) for (int e = 0; e < ne; ++e) {
          const real sgn = e ? real(-1) : real(1);
          real sp[3] = {p2[0] + sgn * hl * a2[0], p2[1] + sgn * hl * a2[1], p2[2] + sgn * hl * a2[2]}, dd, nn[3];
          if (hfield_sphere(cm, E.hfield, p1, sp, r, margin, &dd, nn)) {
            cdst[n] = dd; cn[n][0] = nn[0]; cn[n][1] = nn[1]; cn[n][2] = nn[2];
            const real s = dd * real(0.5) + r;
            cp[n][0] = sp[0] - nn[0] * s; cp[n][1] = sp[1] - nn[1] * s; cp[n][2] = sp[2] - nn[2] * s; ++n;
          }
        }
        if (kind == PAIR_HFIELD_CAPSULE) { L(ch0) = a2[0]; L(ch1) = a2[1]; L(ch2) = a2[2]; }
      } else if ((FEAT & F_BOX) && kind >= PAIR_PLANE_BOX) {
        n = box_pair(kind, margin, p1, p2, cm.geom_size[g1], cm.geom_size[g2], cp, cn, cdst);
        if (kind == PAIR_CAPSULE_BOX && n) { L(ch0) = a1[0]; L(ch1) = a1[1]; L(ch2) = a1[2]; }
      } else if (kind == PAIR_CAPSULE_CAPSULE) {
        const real s1 = cm.geom_size[g1][1], s2 = cm.geom_size[g2][1], r1 = cm.geom_size[g1][0], r2 = cm.geom_size[g2][0];
        real dif[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
        const real ma = dot3(a1, a1), mb = -dot3(a1, a2), mc = dot3(a2, a2), u = -dot3(a1, dif), v = dot3(a2, dif), det = ma * mc - mb * mb;
        real x1c[2], x2c[2]; int nc = 0;
        if (mabs(det) >= minval<real>()) {
          real x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
          if (x1 > s1) { x1 = s1; x2 = (v - mb * s1) / mc; } else if (x1 < -s1) { x1 = -s1; x2 = (v + mb * s1) / mc; }
          if (x2 > s2) { x2 = s2; x1 = clampr((u - mb * s2) / ma, -s1, s1); } else if (x2 < -s2) { x2 = -s2; x1 = clampr((u + mb * s2) / ma, -s1, s1); }
          x1c[0] = x1; x2c[0] = x2; nc = 1;
        } else {  // parallel axes: segment ends of 1 against 2 (at most two contacts are kept)
          x1c[0] = s1; x2c[0] = clampr((v - mb * s1) / mc, -s2, s2); x1c[1] = -s1; x2c[1] = clampr((v + mb * s1) / mc, -s2, s2); nc = 2;
        }
        for (int e = 0; e < nc; ++e) {
          real v1[3] = {p1[0] + a1[0] * x1c[e], p1[1] + a1[1] * x1c[e], p1[2] + a1[2] * x1c[e]}, v2[3] = {p2[0] + a2[0] * x2c[e], p2[1] + a2[1] * x2c[e], p2[2] + a2[2] * x2c[e]};
          real dd[3] = {v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2]}; const real cdist = msqrt(dot3(dd, dd));
          if (cdist <= margin + r1 + r2 && cdist - r1 - r2 < margin) {
            cdst[n] = cdist - r1 - r2;
            if (cdist < minval<real>()) { cn[n][0] = 1; cn[n][1] = cn[n][2] = 0; } else { cn[n][0] = dd[0] / cdist; cn[n][1] = dd[1] / cdist; cn[n][2] = dd[2] / cdist; }
            const real s = r1 + cdst[n] * real(0.5);
            cp[n][0] = v1[0] + cn[n][0] * s; cp[n][1] = v1[1] + cn[n][1] * s; cp[n][2] = v1[2] + cn[n][2] * s; ++n;
          }
        }
      }
      L(ccount) = n;
      for (int e = 0; e < n; ++e) { LA(cb, 7 * e) = cp[e][0]; LA(cb, 7 * e + 1) = cp[e][1]; LA(cb, 7 * e + 2) = cp[e][2]; LA(cb, 7 * e + 3) = cn[e][0]; LA(cb, 7 * e + 4) = cn[e][1]; LA(cb, 7 * e + 5) = cn[e][2]; LA(cb, 7 * e + 6) = cdst[e]; }
    }
    L(coff) = L(ccount);
  ENDL
  int pass_total = 0;
  EXSCAN_INT(coff, pass_total);
  LANES  // write contacts in pair order: [pos3 frame9 dist pair]
    for (int e = 0; e < L(ccount); ++e) {
      const int c = ncon_total + L(coff) + e;
      if (c < MAXCON) {
        real *o = con + 16 * c;
        o[0] = LA(cb, 7 * e); o[1] = LA(cb, 7 * e + 1); o[2] = LA(cb, 7 * e + 2);
        real f[9] = {LA(cb, 7 * e + 3), LA(cb, 7 * e + 4), LA(cb, 7 * e + 5), L(ch0), L(ch1), L(ch2), 0, 0, 0};
        make_frame(f);
        for (int k = 0; k < 9; ++k) o[3 + k] = f[k];
        o[12] = LA(cb, 7 * e + 6); o[13] = (real)(pbase + l); o[15] = (real)(8 * pair_rank(cm.pair_code[pbase + l]) + e);   // pair, (rank, e) sort key
      }
    }
  ENDL
  ncon_total += pass_total;
  }
  }
  if ((FEAT & F_BOX) && ncon_total > ncon_a && ncon_a > 0) {   // contacts of both runs: back into MuJoCo's order (by pair rank, then contact index within the pair)
    const int nc = ncon_total < MAXCON ? ncon_total : MAXCON;
    LVA(real, row, 16); LV(int, dst);
    LANES
      L(dst) = -1;
      if (l < nc) { const real key = con[16 * l + 15]; int r = 0; for (int c = 0; c < nc; ++c) if (con[16 * c + 15] < key) ++r; L(dst) = r; for (int k = 0; k < 16; ++k) LA(row, k) = con[16 * l + k]; }
    ENDL
    LANES if (L(dst) >= 0) for (int k = 0; k < 16; ++k) con[16 * L(dst) + k] = LA(row, k); ENDL
  }
  int ncon = ncon_total < MAXCON ? ncon_total : MAXCON;

  // ================= constraint rows: J (into Y), pos, source; order = equality, limits, contacts =================
  int nefc = 0;
  for (int e = 0; e < cm.neq; ++e) {
    const int b1 = cm.eq_b1[e], b2 = cm.eq_b2[e];
    LANES  // lane = dof
      real p1[3], p2[3], v[3], j1[3], j2[3], cd[6] = {L(cd0), L(cd1), L(cd2), L(cd3), L(cd4), L(cd5)}, cm3[3] = {L(com0), L(com1), L(com2)};
      mat_vec(v, xmat + 9 * b1, cm.eq_data[e]); p1[0] = xpos[3 * b1] + v[0]; p1[1] = xpos[3 * b1 + 1] + v[1]; p1[2] = xpos[3 * b1 + 2] + v[2];
      mat_vec(v, xmat + 9 * b2, cm.eq_data[e] + 3); p2[0] = xpos[3 * b2] + v[0]; p2[1] = xpos[3 * b2 + 1] + v[1]; p2[2] = xpos[3 * b2 + 2] + v[2];
      jac_col(cm, l, b1, cd, p1, cm3, j1); jac_col(cm, l, b2, cd, p2, cm3, j2);
      for (int k = 0; k < 3; ++k) {
        Y[(nefc + k) * ys + l] = j1[k] - j2[k];
        if (xb >= 0 && l < 6) Y[(nefc + k) * ys + 32 + l] = 0;
        if (l == 0) { efc[4 * (nefc + k) + E_POS] = p1[k] - p2[k]; efc[4 * (nefc + k) + E_SRC] = (real)e; efc[4 * (nefc + k) + E_INEQ] = 0; efc[4 * (nefc + k) + E_SIDE] = (real)side_of(cm.body_side[b1], cm.body_side[b2]); }
      }
    ENDL
    nefc += 3;
  }
  LV(int, lcount); LV(int, loff);
  LANES  // lane = joint: violated limits
    L(lcount) = 0;
    if (l < cm.njnt && cm.jnt_limited[l]) { const real q = qpos[cm.jnt_qposadr[l]]; L(lcount) = (q - cm.jnt_range[l][0] < 0 ? 1 : 0) + (cm.jnt_range[l][1] - q < 0 ? 1 : 0); }
    L(loff) = L(lcount);
  ENDL
  int nlim = 0;
  EXSCAN_INT(loff, nlim);
  LANES
    if (L(lcount) > 0) {
      const real q = qpos[cm.jnt_qposadr[l]]; int r = nefc + L(loff);
      for (int side = 0; side < 2; ++side) {
        const real dist = side ? cm.jnt_range[l][1] - q : q - cm.jnt_range[l][0];
        if (dist < 0 && r < NEFC) {
          for (int d = 0; d < ys - 1; ++d) Y[r * ys + d] = 0;
          Y[r * ys + cm.jnt_dofadr[l]] = side ? real(-1) : real(1);
          efc[4 * r + E_POS] = dist; efc[4 * r + E_SRC] = (real)(64 + l); efc[4 * r + E_INEQ] = 1; efc[4 * r + E_SIDE] = (real)cm.body_side[cm.jnt_body[l]]; ++r;
        }
      }
    }
  ENDL
  nefc += nlim; if (nefc > NEFC) nefc = NEFC;
  const int nefc_before_contacts = nefc;
  int ncon_used = 0;
  for (int c = 0; c < ncon; ++c) {
    const real *o = con + 16 * c; const int p = (int)o[13]; const uint32_t pcode = cm.pair_code[p]; const int rows = cm.pc_condim[pair_pc(pcode)] > 1 ? 4 : 1;
    if (nefc + rows > NEFC) break;
    const int b1 = cm.geom_body[pair_g1(pcode)], b2 = cm.geom_body[pair_g2(pcode)];
    LANES  // lane = dof
      real j1[3], j2[3], cd[6] = {L(cd0), L(cd1), L(cd2), L(cd3), L(cd4), L(cd5)}, cm3[3] = {L(com0), L(com1), L(com2)};
      jac_col(cm, l, b1, cd, o, cm3, j1); jac_col(cm, l, b2, cd, o, cm3, j2);
      real dj[3] = {j2[0] - j1[0], j2[1] - j1[1], j2[2] - j1[2]};
      const real jn = dot3(o + 3, dj);
      if (rows == 1) Y[nefc * ys + l] = jn;
      else {
        const real mu = pair_friction(cm, ce, p), jt1 = dot3(o + 6, dj), jt2 = dot3(o + 9, dj);
        Y[nefc * ys + l] = jn + mu * jt1; Y[(nefc + 1) * ys + l] = jn - mu * jt1;
        Y[(nefc + 2) * ys + l] = jn + mu * jt2; Y[(nefc + 3) * ys + l] = jn - mu * jt2;
      }
      if (xb >= 0 && l < 6) {   // columns of the extra free body's dofs: translation = world axes, rotation = body axes about its origin
        real col[3] = {0, 0, 0};
        if (b1 == xb || b2 == xb) {
          if (l < 3) col[l] = 1;
          else { const real axv[3] = {xmat[9 * xb + (l - 3)], xmat[9 * xb + 3 + (l - 3)], xmat[9 * xb + 6 + (l - 3)]}, off[3] = {o[0] - xpos[3 * xb], o[1] - xpos[3 * xb + 1], o[2] - xpos[3 * xb + 2]}; cross3(col, axv, off); }
          if (b1 == xb) { col[0] = -col[0]; col[1] = -col[1]; col[2] = -col[2]; }
        }
        const real xn = dot3(o + 3, col);
        if (rows == 1) Y[nefc * ys + 32 + l] = xn;
        else {
          const real mu = pair_friction(cm, ce, p), xt1 = dot3(o + 6, col), xt2 = dot3(o + 9, col);
          Y[nefc * ys + 32 + l] = xn + mu * xt1; Y[(nefc + 1) * ys + 32 + l] = xn - mu * xt1;
          Y[(nefc + 2) * ys + 32 + l] = xn + mu * xt2; Y[(nefc + 3) * ys + 32 + l] = xn - mu * xt2;
        }
      }
      if (l < rows) { efc[4 * (nefc + l) + E_POS] = o[12]; efc[4 * (nefc + l) + E_SRC] = (real)(128 + p); efc[4 * (nefc + l) + E_INEQ] = 1; efc[4 * (nefc + l) + E_SIDE] = (real)side_of(cm.body_side[b1], cm.body_side[b2]); }
      if (l == 0) con[16 * c + 14] = (real)nefc;   // first row of this contact (read back by the contact-force stage)
    ENDL
    nefc += rows; ++ncon_used;
  }
  if (counters) { LANES if (l == 0) { counters[0] = nefc; counters[1] = ncon_used; counters[2] = nlim; if (ncon_total > ncon_used) counters[4] += ncon_total - ncon_used; } ENDL }
  (void)nefc_before_contacts;
  // ================= derived quantities, part 2 (while the kinematics buffers are alive): foot positions, toe / heel points, a copy of
  // cdof for the foot velocities (cassie_sim_foot_positions :1608-1621; site_xpos of the toe / heel sites :1888-1889)
  if (auxr) {
    real *cdofs = sm + scratch_reals(ys);   // the extended instance's tail
    LANES
      if (l < nv) for (int k = 0; k < 6; ++k) cdofs[6 * l + k] = cdof[6 * l + k];
      if (l < 6) { const int fb = cm.foot_body[l / 3], k = l % 3; real v = fb >= 0 ? xpos[3 * fb + k] : real(0); if (k == 2) v -= cm.foot_offset; auxr[AX_FOOT_POS + l] = v; }
      if (l >= 8 && l < 16) {
        const int i = l - 8, sd = i >> 2, k = i & 1, fb = cm.foot_body[sd]; const real *loc = ((i >> 1) & 1) ? cm.heel_local[sd] : cm.toe_local[sd];
        auxr[AX_TMP + i] = fb >= 0 ? xpos[3 * fb + k] + (xmat[9 * fb + 3 * k] * loc[0] + xmat[9 * fb + 3 * k + 1] * loc[1] + xmat[9 * fb + 3 * k + 2] * loc[2]) : real(0);
      }
    ENDL
  }

  STAGE_SYNC(csync & 4);
  LV(real, qacc); LV(real, qfrc_con);
  LV(real, f0); LV(real, f1);    // constraint forces: lane (r & 31) owns rows r and r + 32
  int iters = 0;
  if (nefc == 0) {
    LANES L(qacc) = L(qacc_smooth); L(qfrc_con) = 0; L(f0) = L(f1) = 0; L(xqacc) = L(xqacc_smooth); ENDL
  } else {
    // ---- per-row: impedance, R, aref, b, warm-start force (lane = row)
    LANES if (l < nv) { vecs[l] = L(qvel); vecs[32 + l] = L(qacc_smooth); vecs[64 + l] = L(qacc_ws); } L(f0) = L(f1) = 0;
      if (xb >= 0 && l < 6) { vecs[160 + l] = L(xqvel); vecs[166 + l] = L(xqacc_smooth); vecs[172 + l] = L(xqacc_ws); } ENDL
    if (dbg) { LANES for (int r = 0; r < nefc; ++r) dbg[D_J + 32 * r + l] = Y[r * ys + l]; ENDL }
    LV(int, side);   // lane r: which leg row r touches (E_SIDE), kept for the products below
    LANES L(side) = 0; ENDL
    for (int pass = 0; pass * 32 < nefc; ++pass) {
      // rows that touch a single leg take the mirrored-leg transform; one row over both legs (leg-leg contact) sends the whole pass the long way
      LV(int, anyb);
      LANES { const int r = l + 32 * pass; L(anyb) = (r < nefc && (int)efc[4 * r + E_SIDE] == 3) ? 1 : 0; } ENDL
      ALLMAX(anyb);
      const bool symrows = cm.sym_on && LANE0(anyb) == 0;
      LANES
        const int r = l + 32 * pass;
        if (r < nefc) {
          const int sd = (int)efc[4 * r + E_SIDE]; if (pass == 0) L(side) = symrows ? sd : 3;
          real *yy = Y + r * ys; real jv = 0, ja = 0, jw = 0;
          // single-leg rows on the tree csrc/cassie_tree_gen.inc was generated for: the row's 19 possibly non-zero entries (base chain + its leg) are held in
          // registers from here to the end of the transform; the other leg's entries are exact zeros and stay untouched in shared memory
          const bool spec = symrows && cm.spec19;
          const int loff = sd == 2 ? 13 : 0;
          real y19[19];
          if (spec) {
#pragma unroll
            for (int d = 0; d < 6; ++d) y19[d] = yy[d];
#pragma unroll
            for (int d = 0; d < 13; ++d) y19[6 + d] = yy[6 + loff + d];
#pragma unroll
            for (int d = 0; d < 19; ++d) { const int dd = d < 6 ? d : d + loff; const real y = y19[d]; jv += y * vecs[dd]; ja += y * vecs[32 + dd]; jw += y * vecs[64 + dd]; }
          } else
          for (int d = 0; d < nv; ++d) { const real y = yy[d]; jv += y * vecs[d]; ja += y * vecs[32 + d]; jw += y * vecs[64 + d]; }
          if (xb >= 0) for (int d = 0; d < 6; ++d) { const real y = yy[32 + d]; jv += y * vecs[160 + d]; ja += y * vecs[166 + d]; jw += y * vecs[172 + d]; }
          const int src = (int)efc[4 * r + E_SRC]; const real pos = efc[4 * r + E_POS]; const bool ineq = efc[4 * r + E_INEQ] != 0;
          const real *solref, *solimp; real dA, margin = 0, rscale = 1;
          if (src < 64) { solref = cm.eq_solref[src]; solimp = cm.eq_solimp[src]; dA = binvw[cm.eq_b1[src]] + binvw[cm.eq_b2[src]]; }
          else if (src < 128) { const int j = src - 64; solref = cm.jnt_solref[j]; solimp = cm.jnt_solimp[j]; dA = dinvw[cm.jnt_dofadr[j]]; }
          else {
            const int p = src - 128; const uint32_t pcode = cm.pair_code[p]; const int pc = pair_pc(pcode);
            solref = cm.pc_solref[pc]; solimp = cm.pc_solimp[pc]; margin = cm.pc_margin[pc] - cm.pc_gap[pc];
            dA = binvw[cm.geom_body[pair_g1(pcode)]] + binvw[cm.geom_body[pair_g2(pcode)]];
            if (cm.pc_condim[pc] > 1) { const real mu = pair_friction(cm, ce, p); dA += mu * mu * dA; rscale = 2 * mu * mu; }
          }
          const real imp = impedance(solimp, pos, margin);
          const real Rr = rscale * mmax(minval<real>(), (1 - imp) * dA / imp);
          real K, B;
          if (solref[0] > 0) { const real tc = mmax(solref[0], 2 * cm.timestep); K = 1 / mmax(minval<real>(), solimp[1] * solimp[1] * tc * tc * solref[1] * solref[1]); B = 2 / mmax(minval<real>(), solimp[1] * tc); }
          else { K = -solref[0] / mmax(minval<real>(), solimp[1] * solimp[1]); B = -solref[1] / mmax(minval<real>(), solimp[1]); }
          const real aref = -B * jv - K * imp * (pos - margin);
          const real jar = jw - aref; real f = -jar / Rr;
          if (ineq && jar >= 0) f = 0;
          if (pass == 0) L(f0) = f; else L(f1) = f;
          if (dbg) { dbg[D_EFC_AREF + r] = aref; dbg[D_EFC_R + r] = Rr; dbg[D_EFC_B + r] = ja - aref; }
          // ---- Y row <- sqrt(inv(D)) inv(L') J row  (mj_solveM2), in place; all lanes (rows) walk the same (i, ancestor) sequence
          real ad;
          if (spec) {
            const real *qLb = sm + S_QLD, *qLl = sm + S_QLD + (sd == 2 ? cm.sym_madr : 0);
#define Y19_(i) y19[i]
#define LB19_(a) qLb[a]
#define LL19_(a) qLl[a]
            CT19_TRANSFORM(Y19_, LB19_, LL19_, real);
#undef Y19_
#undef LB19_
#undef LL19_
            ad = 0;
#pragma unroll
            for (int d = 0; d < 19; ++d) { const int dd = d < 6 ? d : d + loff; const real v = y19[d] * sm[S_DSQI + dd]; yy[dd] = v; ad += v * v; }
          } else ad = symrows ? half_solve_row_sym(cm, sm, yy, sd == 2 ? cm.sym_n : 0) : half_solve_row(cm, sm, yy, nv);
          if (xb >= 0) for (int d = 0; d < 6; ++d) { const real v = yy[32 + d] * (d < 3 ? xb_dsqi_t : cm.xb_dsqi[d]); yy[32 + d] = v; ad += v * v; }
          // packed row constants for the solver: b, 1/A, A, +-R (sign bit set = inequality row)
          real *rc = efc + 4 * r; const real Ad = ad + Rr;
          rc[0] = ja - aref; rc[1] = real(1) / Ad; rc[2] = Ad; rc[3] = ineq ? -Rr : Rr;
        }
      ENDL
    }
    LV(real, z); LV(real, fb); LV(real, impr);
    if (nefc <= 32 && !cm.force_zpath) {
      // ---- dense path (nefc <= 32: standing / walking contact states): A = Y Y' + R is formed explicitly in scratch that is idle during the
      // solve, and the solver carries the residual res = b + A f one entry per lane: a Gauss-Seidel row update is two
      // broadcasts, ~10 scalar ops, one shared load and one FMA per lane, with no warp reduction on the critical path
      // A(., c) for c < 16 lives in rows 32..47 of the Y region, for c >= 16 in the (now dead) kinematics buffers xpos/xquat/xmat/cdof
#define AM(c) (((c) < 16 ? Y + (32 + (c)) * ys : sm + S_XPOS + ((c) - 16) * ys))
      {
        LVA(real, yreg, 32);
        LANES_NS
#pragma unroll
          for (int d = 0; d < 32; ++d) LA(yreg, d) = (l < nefc) ? Y[l * ys + d] : real(0);
        ENDL_NS
        LV(int, sc);
        for (int c = 0; c < nefc; ++c) {
          BCAST(sc, side, c);   // column c touches the base and one leg only: its entries on the other leg are exact zeros
          LANES_NS
            const real *yc = Y + c * ys; real sacc = 0;
            if (LANE0(sc) == 3) {
#pragma unroll
              for (int d = 0; d < 32; ++d) sacc += LA(yreg, d) * yc[d];
            } else if (LANE0(sc) == 2) {
#pragma unroll
              for (int d = 0; d < 6; ++d) sacc += LA(yreg, d) * yc[d];
#pragma unroll
              for (int d = 19; d < 32; ++d) sacc += LA(yreg, d) * yc[d];
            } else {
#pragma unroll
              for (int d = 0; d < 19; ++d) sacc += LA(yreg, d) * yc[d];
            }
            if (xb >= 0 && l < nefc) for (int d = 32; d < 38; ++d) sacc += Y[l * ys + d] * yc[d];
            if (c == l) sacc += mabs(efc[4 * c + 3]);
            AM(c)[l] = sacc;
          ENDL_NS
        }
      }
      // warm start: res = b + A f; keep f only if its dual cost f'(b + 0.5 A f) is not positive
      LV(real, res); LV(real, ri);
      LANES L(res) = 0; ENDL
      for (int c = 0; c < nefc; ++c) { BCAST(fb, f0, c); LANES_NS L(res) += AM(c)[l] * L(fb); ENDL_NS }
      LANES_NS
        const real bl = (l < nefc) ? efc[4 * l] : real(0);
        L(t0) = (l < nefc) ? L(f0) * (bl + real(0.5) * L(res)) : real(0);
        L(res) += bl; L(t1) = bl;
      ENDL_NS
      ALLSUM(t0);
      if (LANE0(t0) > 0) { LANES_NS L(f0) = 0; L(res) = L(t1); ENDL_NS }
      // solver constants repacked for the sweep: slot 2 <- A_ii / 2, slot 3 <- the row's lower bound (0: inequality, -inf: equality)
      LANES_NS if (l < nefc) { real *rc = efc + 4 * l; rc[2] = real(0.5) * rc[2]; rc[3] = rc[3] < 0 ? real(0) : -(real)INFINITY; } ENDL_NS
      // One row update = a scalar chain every lane runs redundantly (f_i, its change, the cost change) + one FMA per lane (res += A(.,i) delta).  The
      // residual of the NEXT row is not fetched from its lane after this row's update (a shuffle on the critical path) but rebuilt by every lane from the
      // value broadcast one row earlier plus this row's contribution -- the same FMA on the same operands as that lane performs itself.
      LV(real, rs); LV(real, fs);   // fs: the forces as they stood at the start of the sweep (row i is the only writer of f_i, so f_i is still that value when row i reads it)
#define PGS_ROW_(i, Ac)                                                                                                              \
  BCAST(rs, res, (i) + 1 < 32 ? (i) + 1 : 31); BCAST(fb, fs, i);                                                                     \
  LANES_NS                                                                                                                           \
    const Row4<real> rc = load_row4(efc + 4 * (i));                                                                                  \
    const real ainv = rc.y, hAd = rc.z, lo = rc.w, fold = L(fb), resi = L(ri);                                                       \
    real fnew = mmax(fold - resi * ainv, lo);                                                                                        \
    real delta = fnew - fold;                                                                                                        \
    const real change = delta * (hAd * delta + resi);                                                                                \
    if (change > real(1e-10)) { delta = 0; fnew = fold; } else L(impr) -= change;                                                    \
    L(res) += (Ac)[l] * delta;                                                                                                       \
    L(ri) = L(rs) + (Ac)[(i) + 1] * delta;                                                                                           \
    if (l == (i)) L(f0) = fnew;                                                                                                      \
  ENDL_NS
      while (iters < cm.iterations) {
        LANES_NS L(impr) = 0; L(fs) = L(f0); ENDL_NS
        BCAST(ri, res, 0);
        {
          const real *Ac = Y + 32 * ys; const int n0 = nefc < 16 ? nefc : 16;
          for (int i = 0; i < n0; ++i, Ac += ys) { PGS_ROW_(i, Ac) }
          Ac = sm + S_XPOS;
          for (int i = 16; i < nefc; ++i, Ac += ys) { PGS_ROW_(i, Ac) }
        }
        ++iters;
        if (LANE0(impr) * pgs_scale < cm.tolerance) break;
      }
#undef PGS_ROW_
      LANES_NS L(z) = 0; ENDL_NS
      for (int r = 0; r < nefc; ++r) { BCAST(fb, f0, r); LANES_NS if (l < nv) L(z) += Y[r * ys + l] * L(fb); if (xb >= 0 && l < 6) L(xz) += Y[r * ys + 32 + l] * L(fb); ENDL_NS }
    } else {
      // ---- warm start: keep f only if its dual cost 0.5 f'(YY'+R)f + f'b is not positive
      LANES L(z) = 0; ENDL
      for (int r = 0; r < nefc; ++r) {
        if (r < 32) { BCAST(fb, f0, r); } else { BCAST(fb, f1, r - 32); }
        LANES if (l < nv) L(z) += Y[r * ys + l] * L(fb); if (xb >= 0 && l < 6) L(xz) += Y[r * ys + 32 + l] * L(fb); ENDL
      }
      LANES
        real s = real(0.5) * (L(z) * L(z) + L(xz) * L(xz));
        if (l < nefc) { const real f = L(f0); s += f * (efc[4 * l] + real(0.5) * mabs(efc[4 * l + 3]) * f); }
        if (l + 32 < nefc) { const real f = L(f1); s += f * (efc[4 * (l + 32)] + real(0.5) * mabs(efc[4 * (l + 32) + 3]) * f); }
        L(t0) = s;
      ENDL
      ALLSUM(t0);
      if (LANE0(t0) > 0) { LANES L(z) = 0; L(xz) = 0; L(f0) = L(f1) = 0; ENDL }
      // ---- projected Gauss-Seidel, rows strictly in order; z = Y'f is carried one entry per lane
      LV(real, acc); LV(real, yr); LV(real, yx);
      while (iters < cm.iterations) {
        LANES L(impr) = 0; ENDL
        const int n0 = nefc < 32 ? nefc : 32;
        for (int r = 0; r < n0; ++r) {
          BCAST(fb, f0, r);
          LANES_NS L(yr) = (l < nv) ? Y[r * ys + l] : real(0); L(yx) = (xb >= 0 && l < 6) ? Y[r * ys + 32 + l] : real(0); L(acc) = L(yr) * L(z) + L(yx) * L(xz); ENDL_NS
          ALLSUM(acc);
          LANES_NS
            const real *rc = efc + 4 * r;
            const real b = rc[0], ainv = rc[1], Ad = rc[2], Rs = rc[3], fold = L(fb);
            const real res = b + L(acc) + mabs(Rs) * fold;
            real fnew = fold - res * ainv;
            if (Rs < 0) fnew = mmax(fnew, real(0));
            real delta = fnew - fold;
            real change = delta * (real(0.5) * delta * Ad + res);
            if (change > real(1e-10)) { delta = 0; change = 0; fnew = fold; }
            L(impr) -= change;
            L(z) += L(yr) * delta; L(xz) += L(yx) * delta;
            if (l == r) L(f0) = fnew;
          ENDL_NS
        }
        for (int r = 32; r < nefc; ++r) {
          BCAST(fb, f1, r - 32);
          LANES_NS L(yr) = (l < nv) ? Y[r * ys + l] : real(0); L(yx) = (xb >= 0 && l < 6) ? Y[r * ys + 32 + l] : real(0); L(acc) = L(yr) * L(z) + L(yx) * L(xz); ENDL_NS
          ALLSUM(acc);
          LANES_NS
            const real *rc = efc + 4 * r;
            const real b = rc[0], ainv = rc[1], Ad = rc[2], Rs = rc[3], fold = L(fb);
            const real res = b + L(acc) + mabs(Rs) * fold;
            real fnew = fold - res * ainv;
            if (Rs < 0) fnew = mmax(fnew, real(0));
            real delta = fnew - fold;
            real change = delta * (real(0.5) * delta * Ad + res);
            if (change > real(1e-10)) { delta = 0; change = 0; fnew = fold; }
            L(impr) -= change;
            L(z) += L(yr) * delta; L(xz) += L(yx) * delta;
            if (l == r - 32) L(f1) = fnew;
          ENDL_NS
        }
        ++iters;
        if (LANE0(impr) * pgs_scale < cm.tolerance) break;
      }
    }
    // ---- qacc = qacc_smooth + inv(L) D^-1/2 z ;  qfrc_constraint = J'f = L' D^1/2 z
    LV(real, w);
    LANES L(w) = (l < nv) ? L(z) * sm[S_DSQI + l] : real(0); ENDL
    solve_l(cm, sm, w);
    LANES L(qacc) = L(qacc_smooth) + L(w); L(qfrc_con) = 0; if (xb >= 0 && l < 6) L(xqacc) = L(xqacc_smooth) + L(xz) * (l < 3 ? xb_dsqi_t : cm.xb_dsqi[l]); ENDL
    if (dbg) {  // qfrc_constraint = M (qacc - qacc_smooth); only the debug dump wants it (the Euler stage below does not)
      LANES vecs[128 + l] = L(w); ENDL
      LANES
        if (l < nv) {
          real sacc = 0; int a = cm.dof_Madr[l];
          for (int j = l; j >= 0; j = cm.dof_parent[j]) sacc += qM[a++] * vecs[128 + j];
          for (int i = l + 1; i < cm.dof_subtree_end[l]; ++i) sacc += qM[cm.dof_Madr[i] + cm.dof_depth[i] - cm.dof_depth[l]] * vecs[128 + i];
          L(qfrc_con) = sacc;
        }
      ENDL
    }
  }
  STAGE_SYNC(csync & 8);
  if (counters) { LANES if (l == 0) counters[3] = iters; ENDL }
  // ================= derived quantities, part 3: contact forces (mj_contactForce -> world frame), foot / toe / heel sums, collision flags
  // (cassie_sim_foot_forces :1812-1854, cassie_sim_heeltoe_forces :1856-1898, check_*_collision :1586-1606, geom_collision :1944-1961)
  if (auxr) {
    real *aux = auxr;
    if (nefc > 0) { LANES if (l < nefc) efc[4 * l] = L(f0); if (l + 32 < nefc) efc[4 * (l + 32)] = L(f1); ENDL }   // slot 0 (b) is dead: keep the row forces
    LANES  // lane = contact: decode the pyramid, rotate into the world frame, classify
      if (l < ncon_used) {
        real *o = con + 16 * l; const int p = (int)o[13], r0 = (int)o[14]; const uint32_t pcode = cm.pair_code[p];
        real fn, t1 = 0, t2 = 0;
        if (cm.pc_condim[pair_pc(pcode)] == 1) fn = efc[4 * r0];
        else { const real a = efc[4 * r0], b = efc[4 * (r0 + 1)], c = efc[4 * (r0 + 2)], d = efc[4 * (r0 + 3)], mu = pair_friction(cm, ce, p); fn = a + b + c + d; t1 = (a - b) * mu; t2 = (c - d) * mu; }
        real F[3]; for (int k = 0; k < 3; ++k) F[k] = o[3 + k] * fn + o[6 + k] * t1 + o[9 + k] * t2;
        const int b1 = cm.geom_body[pair_g1(pcode)], b2 = cm.geom_body[pair_g2(pcode)], lf = cm.foot_body[0], rf = cm.foot_body[1];
        const bool f1 = (b1 == lf || b1 == rf), anyf = f1 || b2 == lf || b2 == rf; const int id = (b1 == rf || b2 == rf) ? 1 : 0;
        real toe = 0;
        if (anyf) {
          const real td0 = aux[AX_TMP + 4 * id] - o[0], td1 = aux[AX_TMP + 4 * id + 1] - o[1], hd0 = aux[AX_TMP + 4 * id + 2] - o[0], hd1 = aux[AX_TMP + 4 * id + 3] - o[1];
          toe = (msqrt(td0 * td0 + td1 * td1) < msqrt(hd0 * hd0 + hd1 * hd1)) ? real(1) : real(0);
        }
        o[3] = F[0]; o[4] = F[1]; o[5] = F[2];
        o[6] = b1 == lf ? real(-1) : (b2 == lf ? real(1) : real(0));   // sign with which this contact enters the left / right foot force
        o[7] = b1 == rf ? real(-1) : (b2 == rf ? real(1) : real(0));
        o[8] = anyf ? (f1 ? real(-1) : real(1)) : real(0); o[9] = (real)id; o[10] = toe;   // heel / toe split: sign, foot, toe-or-heel
      }
    ENDL
    LANES  // lane = output slot
      if (l < 18) {
        real acc = 0;
        for (int c = 0; c < ncon_used; ++c) {
          const real *o = con + 16 * c;
          if (l < 6) { const real sg = o[6 + l / 3]; if (sg != 0) acc += sg * o[3 + l % 3]; }
          else { const int i = l - 6, j = i % 6; if (o[8] != 0 && (int)o[9] == j / 3 && (o[10] != 0) == (i < 6)) acc += o[8] * o[3 + j % 3]; }
        }
        if (l < 6) { aux[AX_FOOT_FORCE + 6 * (l / 3) + l % 3] = acc; aux[AX_FOOT_FORCE + 6 * (l / 3) + 3 + l % 3] = 0; }
        else aux[AX_TOE_FORCE + (l - 6)] = acc;
      } else if (l == 18) {
        int fl = 0; for (int c = 0; c < ncon_used; ++c) fl |= cm.pc_flags[pair_pc(cm.pair_code[(int)con[16 * c + 13]])];
        aux[AX_OBSTACLE] = (fl & 1) ? real(1) : real(0); aux[AX_SELF] = (fl & 2) ? real(1) : real(0); aux[AX_GROUPMASK] = (real)(fl >> 8); aux[AX_NCON] = (real)ncon_used;
      }
    ENDL
  }
  if (dbg) {
    LANES
      if (l < nv) { dbg[D_QACC + l] = L(qacc); dbg[D_QFRCC + l] = L(qfrc_con); }
      if (l < nefc) dbg[D_EFC_F + l] = L(f0);
      if (l + 32 < nefc) dbg[D_EFC_F + l + 32] = L(f1);
      if (l == 0) { dbg[D_COUNTS] = (real)nefc; dbg[D_COUNTS + 1] = (real)ncon_used; dbg[D_COUNTS + 2] = (real)nlim; dbg[D_COUNTS + 3] = (real)iters; }
    ENDL
  }

  real *cst = E.cst;
  // ================= accelerometer (mj_sensorAcc): the position / velocity sensors and the qacc-independent part were prepared before
  // the constraint stage (which recycles the kinematics buffers); accel = g0 + G qacc over the IMU body's dof chain
  LANES if (l < nv) vecs[32 + l] = L(qacc); ENDL
  LANES
    if (l == 16) {
      const int nchain = (int)vecs[127];
      for (int k = 0; k < 3; ++k) {
        real acc = vecs[102 + k]; int a = cm.body_lastdof[cm.imu_body];
        for (int c = 0; c < nchain; ++c, a = cm.dof_parent[a]) acc += vecs[105 + 3 * c + k] * vecs[32 + a];
        if (cm.accel_cutoff > 0) acc = clampr(acc, -cm.accel_cutoff, cm.accel_cutoff);
        cst[CS_SENSOR + 23 + k] = acc;
      }
    }
  ENDL
  if (dbg) { LANES if (l < 29) dbg[D_SENS + l] = cst[CS_SENSOR + l]; ENDL }

  if (!advance) {   // mj_forward only (used once at init / reset to populate sensordata, src/cassiemujoco.c:1029)
    if (auxr) { LANES if (l < nv) vecs[l] = L(qvel); ENDL aux_foot_velocities(cm, sm, auxr); }
    return;
  }
  STAGE_SYNC(csync & 16);
  // ================= Euler with implicit joint damping (mj_Euler) + mj_advance =================
  // (M + hB) a = qfrc_smooth + qfrc_constraint = M qacc   =>   a = qacc - c  with  (M + hB) c = hB qacc  (exact; no J'f needed).
  LV(real, a);
  if (cm.has_damping || ce) {
    LANES   // the factor of M is dead: its buffer takes the factor of M + h B.  All loads first, then the stores: one L2 round trip instead of nM / 32 dependent ones
      real t[NM_MAX / 32];
#pragma unroll
      for (int j = 0; j < NM_MAX / 32; ++j) { const int k = l + 32 * j; t[j] = k < cm.nM ? qM[NM_MAX + k] : real(0); }
#pragma unroll
      for (int j = 0; j < NM_MAX / 32; ++j) { const int k = l + 32 * j; if (k < cm.nM) qLD[k] = t[j]; }
    ENDL
    LANES if (l < nv) sm[S_DINV + l] = mrcp(qLD[cm.dof_Madr[l]]); ENDL
    LANES L(a) = (l < nv) ? cm.timestep * ddamp[l] * L(qacc) : real(0); ENDL
    solve_m(cm, sm, a);
    LANES L(a) = L(qacc) - L(a); ENDL
  } else { LANES L(a) = L(qacc); ENDL }
  LANES if (l < nv) { L(qvel) += cm.timestep * L(a); vecs[l] = L(qvel); L(qacc_ws) = L(qacc); }
    if (xb >= 0 && l < 6) { L(xqvel) += cm.timestep * L(xqacc); vecs[160 + l] = L(xqvel); L(xqacc_ws) = L(xqacc); } ENDL
  if (auxr) aux_foot_velocities(cm, sm, auxr);   // with the NEW qvel and the cdof of this sub-step's kinematics, as the reference's query does
  LANES  // lane = joint: integrate positions with the NEW velocity
    if (l < cm.njnt) {
      const int t = cm.jnt_type[l], qa = cm.jnt_qposadr[l], da = cm.jnt_dofadr[l]; const real h = cm.timestep;
      if (t >= 2) qpos[qa] += h * vecs[da];
      else {
        const real *vv = (t == 1) ? vecs + da : vecs + 160 + 3;     // ball: main-tree velocities; free: the extra body's angular velocity
        const int qq = (t == 1) ? qa : qa + 3;
        if (t == 0) { qpos[qa] += h * vecs[160]; qpos[qa + 1] += h * vecs[161]; qpos[qa + 2] += h * vecs[162]; }
        real wv[3] = {vv[0], vv[1], vv[2]}, q[4] = {qpos[qq], qpos[qq + 1], qpos[qq + 2], qpos[qq + 3]}, qr[4], s, c;
        const real ang = h * normalize3(wv);
        msincos(real(0.5) * ang, &s, &c); qr[0] = c; qr[1] = wv[0] * s; qr[2] = wv[1] * s; qr[3] = wv[2] * s;
        if (ang == 0) { qr[0] = 1; qr[1] = qr[2] = qr[3] = 0; }
        normalize4(q); mul_quat(q, q, qr);
        qpos[qq] = q[0]; qpos[qq + 1] = q[1]; qpos[qq + 2] = q[2]; qpos[qq + 3] = q[3];
      }
    }
    if (l == 0) cst[CS_TIME] += cm.timestep;
  ENDL
}

// rotation matrix (row-major) -> quaternion with MuJoCo's mat2quat branches; the estimator's quaternion signs follow them
template <typename real>
CFN void est_mat2quat(real *q, const real *R) {
  const real t = R[0] + R[4] + R[8];
  if (t > 0) { const real s = msqrt(t + 1) * 2; q[0] = real(0.25) * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { const real s = msqrt(1 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = real(0.25) * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { const real s = msqrt(1 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = real(0.25) * s; q[3] = (R[5] + R[7]) / s; }
  else { const real s = msqrt(1 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = real(0.25) * s; }
}

// ------------------------------------------------------------------ estimator, stateless part (state_output_step [closed], decoded)
// Foot pose and velocity of one leg from the MEASURED angles: chain pelvis -> hipRoll -> hipYaw -> hipPitch, then the planar part knee ->
// shin -> tarsus -> foot (parallel axes: angles add up).  sc = {sin, cos} of hipRoll, hipYaw, hipPitch and of the four cumulative planar
// angles; rate = hipRoll, hipYaw, hipPitch, knee, shin, tarsus, foot rates.  out: position 3, quaternion 4 (pelvis frame), rotational and
// translational velocity (3 + 3) in the reported foot frame.  Offsets: model/cassie.xml:96-152; the foot point and the 40 degree frame
// offset are the estimator's own constants.
template <typename real>
CFN void est_foot(int side, const real *sc, const real *rate, real *out, real *jac = (real *)0) {   // jac (optional): [5][6] columns of [Jw; Jv] for the five motor angles
  const real sg = side ? real(-1) : real(1);
  // A = frame of the hip-pitch link in the pelvis frame (row-major), built from the three quarter-turn link frames and the joint turns
  const real s0 = sc[0], c0 = sc[1], s1 = sc[2], c1 = sc[3], s2 = sc[4], c2 = sc[5];
  // A0 = F0 Rz(a0): columns x, y, z of the hip-roll link;  F0 = [0 0 1; 0 1 0; -1 0 0]
  const real A0[9] = {0, 0, 1, s0, c0, 0, -c0, s0, 0};
  real p[3] = {real(0.021), real(0.135) * sg, 0};
  const real anchor0[3] = {p[0], p[1], p[2]}, axis0[3] = {A0[2], A0[5], A0[8]};
  { const real o1[3] = {0, 0, real(-0.07)}; real v[3]; mat_vec(v, A0, o1); p[0] += v[0]; p[1] += v[1]; p[2] += v[2]; }
  // A1 = A0 F1 Rz(a1),  F1 = [0 0 -1; 0 1 0; 1 0 0]  ->  (A0 F1) columns = (A0z, A0y, -A0x)
  real B[9], A1[9];
  for (int r = 0; r < 3; ++r) { B[3 * r] = A0[3 * r + 2]; B[3 * r + 1] = A0[3 * r + 1]; B[3 * r + 2] = -A0[3 * r]; }
  for (int r = 0; r < 3; ++r) { A1[3 * r] = B[3 * r] * c1 + B[3 * r + 1] * s1; A1[3 * r + 1] = -B[3 * r] * s1 + B[3 * r + 1] * c1; A1[3 * r + 2] = B[3 * r + 2]; }
  const real anchor1[3] = {p[0], p[1], p[2]}, axis1[3] = {A1[2], A1[5], A1[8]};
  { const real o2[3] = {0, 0, real(-0.09)}; real v[3]; mat_vec(v, A1, o2); p[0] += v[0]; p[1] += v[1]; p[2] += v[2]; }
  // A2 = A1 F2 Rz(a2),  F2 = [0 1 0; 0 0 -1; -1 0 0]  ->  (A1 F2) columns = (-A1z, A1x, -A1y)
  real A2[9];
  for (int r = 0; r < 3; ++r) { B[3 * r] = -A1[3 * r + 2]; B[3 * r + 1] = A1[3 * r]; B[3 * r + 2] = -A1[3 * r + 1]; }
  for (int r = 0; r < 3; ++r) { A2[3 * r] = B[3 * r] * c2 + B[3 * r + 1] * s2; A2[3 * r + 1] = -B[3 * r] * s2 + B[3 * r + 1] * c2; A2[3 * r + 2] = B[3 * r + 2]; }
  // planar part, in the hip-pitch frame (joint axes = its z): anchors of knee, shin, tarsus, foot and the foot point
  const real offx[5] = {real(0.12), real(0.06068), real(0.43476), real(0.408), real(0.01762)}, offy[5] = {0, real(0.04741), real(0.02), real(-0.04), real(0.05219)};
  real ax[5], ay[5];   // anchor of joint 3..6 and finally the foot point
  ax[0] = offx[0]; ay[0] = offy[0];
  for (int i = 1; i < 5; ++i) { const real s = sc[6 + 2 * (i - 1)], c = sc[7 + 2 * (i - 1)]; ax[i] = ax[i - 1] + c * offx[i] - s * offy[i]; ay[i] = ay[i - 1] + s * offx[i] + c * offy[i]; }
  const real u[3] = {ax[4], ay[4], real(0.0045) * sg};
  real v[3]; mat_vec(v, A2, u);
  const real pos[3] = {p[0] + v[0], p[1] + v[1], p[2] + v[2]};
  // foot frame = A2 Rz(theta) Roff,  Roff = [-c40 0 -s40; s40 0 -c40; 0 -1 0]
  const real st = sc[12], ct = sc[13], c40 = real(0.766044443118978), s40 = real(0.6427876096865393);
  real Rl[9], Rf[9];
  for (int r = 0; r < 3; ++r) { Rl[3 * r] = A2[3 * r] * ct + A2[3 * r + 1] * st; Rl[3 * r + 1] = -A2[3 * r] * st + A2[3 * r + 1] * ct; Rl[3 * r + 2] = A2[3 * r + 2]; }
  for (int r = 0; r < 3; ++r) { Rf[3 * r] = -Rl[3 * r] * c40 + Rl[3 * r + 1] * s40; Rf[3 * r + 1] = -Rl[3 * r + 2]; Rf[3 * r + 2] = -Rl[3 * r] * s40 - Rl[3 * r + 1] * c40; }
  // velocities: hip roll / yaw about their own axes, everything below about the common z of the hip-pitch frame
  real w[3], vl[3], d[3], cr[3];
  const real rz = rate[2] + rate[3] + rate[4] + rate[5] + rate[6];
  for (int k = 0; k < 3; ++k) w[k] = axis0[k] * rate[0] + axis1[k] * rate[1] + A2[3 * k + 2] * rz;
  d[0] = pos[0] - anchor0[0]; d[1] = pos[1] - anchor0[1]; d[2] = pos[2] - anchor0[2]; cross3(cr, axis0, d);
  for (int k = 0; k < 3; ++k) vl[k] = cr[k] * rate[0];
  d[0] = pos[0] - anchor1[0]; d[1] = pos[1] - anchor1[1]; d[2] = pos[2] - anchor1[2]; cross3(cr, axis1, d);
  for (int k = 0; k < 3; ++k) vl[k] += cr[k] * rate[1];
  // planar: z x (u - a_i) = (-(u - a_i)_y, (u - a_i)_x, 0); the hip-pitch joint's anchor is the frame origin
  real px = -(u[1]) * rate[2], py = (u[0]) * rate[2];
  for (int i = 0; i < 4; ++i) { px -= (u[1] - ay[i]) * rate[3 + i]; py += (u[0] - ax[i]) * rate[3 + i]; }
  const real pl[3] = {px, py, 0}; mat_vec(v, A2, pl);
  vl[0] += v[0]; vl[1] += v[1]; vl[2] += v[2];
  out[0] = pos[0]; out[1] = pos[1]; out[2] = pos[2];
  est_mat2quat(out + 3, Rf);
  matT_vec(out + 7, Rf, w); matT_vec(out + 10, Rf, vl);
  if (jac) {   // pelvis-frame Jacobian of the foot point: hip roll, hip yaw, then hip pitch / knee / foot about the common z of the hip-pitch frame
    d[0] = pos[0] - anchor0[0]; d[1] = pos[1] - anchor0[1]; d[2] = pos[2] - anchor0[2]; cross3(cr, axis0, d);
    for (int k = 0; k < 3; ++k) { jac[k] = axis0[k]; jac[3 + k] = cr[k]; }
    d[0] = pos[0] - anchor1[0]; d[1] = pos[1] - anchor1[1]; d[2] = pos[2] - anchor1[2]; cross3(cr, axis1, d);
    for (int k = 0; k < 3; ++k) { jac[6 + k] = axis1[k]; jac[9 + k] = cr[k]; }
    const real bx[3] = {0, ax[0], ax[3]}, by[3] = {0, ay[0], ay[3]};   // anchors of hip pitch (frame origin), knee, foot in the plane
    for (int j = 0; j < 3; ++j) {
      const real pj[3] = {-(u[1] - by[j]), u[0] - bx[j], 0}; mat_vec(v, A2, pj);
      for (int k = 0; k < 3; ++k) { jac[12 + 6 * j + k] = A2[3 * k + 2]; jac[15 + 6 * j + k] = v[k]; }
    }
  }
}

// ------------------------------------------------------------------ estimator, force model + filters (state_output_step [closed], decoded)
// The arithmetic is csrc/estimator_host.h's (the host-side checker runs those functions one environment per call); here the three per-axis Kalman
// filters are spread over the warp: filter f = lane / 8 (x, y, z), index j = lane % 8 inside it, every covariance / gain entry is produced by one
// lane with the host code's own term order.  All filter data sits in shared memory (the constraint-matrix region is idle between ticks) as
// doubles in every batch precision -- the block's unsymmetrised covariance recursion needs them.  Called once per 2 kHz control tick.
// ds layout (doubles): [0..121] the environment's estimator row (devmodel.h ES_*), [128 + 64 f ..] per-filter parameters a1[6] zm[4] Qd[6] misc,
// [320 + 96 f ..] per-filter work: G[N][K] (24), HP[K][N] (24), S[K][2K] (32), (gain reuses G's slot after S is inverted: Kg[N][K] at +80, 16.. no: see offsets)
template <typename real>
CNOINLINE void est_stage(real *sm, const real *cst, real *obs, double *est, const real *sc) {   // sc: {sin, cos} of the 2 x 7 chain angles (foot-pose stage of this tick)
  DECL_LANE
  double *ds = reinterpret_cast<double *>(sm + S_Y);
  real *eo = obs + OB_EST_OUT;
  constexpr int PAR = 128, WRK = 320;   // per-filter parameter block (64 doubles each), per-filter work block (112 doubles each)
  constexpr int W_G = 0, W_HP = 24, W_S = 48, W_KG = 80;
  // ---- leg forces (lanes 0, 1), in the batch precision: the archive evaluates this block in single precision itself
  LANES
    if (l < 2) {
      const int sd = l; const real *mp = cst + CS_DPOS + 5 * sd, *jp = cst + CS_JPOS + 3 * sd, *q = cst + CS_SENSOR + 16;
      const real ang[7] = {mp[0], mp[1], mp[2], mp[3], jp[0], jp[1], mp[4]}, qd[4] = {q[0], q[1], q[2], q[3]};
      real f[3]; estimator_leg_force_t<real>(sd, ang, qd, f, sc + 14 * sd);
      for (int k = 0; k < 3; ++k) { est[ES_FORCE + 3 * sd + k] = (double)f[k]; eo[EO_TOE + 3 * sd + k] = f[k]; }
    }
    // the environment's filter row -> shared memory (the force slots are rewritten above, read back from `est` below)
    for (int i = l; i < ES_FORCE; i += 32) ds[i] = est[i];
  ENDL
  // ---- per-filter scalars (lane 0 of each filter): measurements, contact terms, state prediction, Jacobian row a1
  LANES
    const int f = l >> 3, j = l & 7;
    if (f < 3 && j == 0) {
      const real *q = obs + OB_EST_QUAT, *pl = obs + OB_FOOT, *pr = obs + OB_FOOT + 13, *ac = obs + OB_EST_ACC;
      const double w = q[0], x = q[1], y = q[2], z = q[3];
      const double r0 = f == 0 ? w * w + x * x - y * y - z * z : f == 1 ? 2 * (x * y + w * z) : 2 * (x * z - w * y);
      const double r1 = f == 0 ? 2 * (x * y - w * z) : f == 1 ? w * w - x * x + y * y - z * z : 2 * (y * z + w * x);
      const double r2 = f == 0 ? 2 * (x * z + w * y) : f == 1 ? 2 * (y * z - w * x) : w * w - x * x - y * y + z * z;
      const double yL = -(pl[0] * r0 + pl[1] * r1 + pl[2] * r2), yR = -(pr[0] * r0 + pr[1] * r1 + pr[2] * r2), aw = ac[0] * r0 + ac[1] * r1 + ac[2] * r2;
      const EstimatorContact k = estimator_contact(est[ES_FORCE + 2] + est[ES_FORCE + 2], est[ES_FORCE + 5] + est[ES_FORCE + 5]);   // toeForce + heelForce
      const bool started = ds[0] != 0;
      double *par = ds + PAR + 64 * f, *a1 = par, *zm = par + 6, *Qd = par + 10;
      const double dt = EST_DT, c = EST_DT * EST_GRAV;
      if (f < 2) {
        double *xs = ds + ES_X + 42 * f, *P = xs + 6;
        if (!started) estimator_axis_start_xy(xs, P, yL, yR);
        const double p0 = xs[0], v0 = xs[1], wt = xs[4];
        zm[0] = yL; zm[1] = yR; zm[2] = k.wm; zm[3] = xs[1] + dt * aw;
        Qd[0] = 1e-8; Qd[1] = 1e-8; Qd[2] = k.qL; Qd[3] = k.qR; Qd[4] = 1e-5; Qd[5] = 1e-2;
        a1[0] = 0; a1[1] = 1; a1[2] = 0; a1[3] = 0; a1[4] = 0; a1[5] = 0;
        if (k.contact) { a1[0] = c; a1[2] = -c * wt; a1[3] = -c * (1 - wt); a1[4] = -c * (xs[2] - xs[3]); a1[5] = dt / EST_MASS;
                         xs[1] = v0 + c * (p0 - wt * xs[2] - (1 - wt) * xs[3]) + dt / EST_MASS * xs[5]; }
        xs[0] = p0 + dt * v0;
      } else {
        double *zs = ds + ES_Z, *P = ds + ES_PZ;
        if (!started) { estimator_axis_start_z(zs, P, yL, yR); ds[ES_TERRAIN] = 0; }
        const double p0 = zs[0], v0 = zs[1];
        zm[0] = yL; zm[1] = yR; Qd[0] = 1e-8; Qd[1] = 1e-8; Qd[2] = k.qL; Qd[3] = k.qR; Qd[4] = 1e-2;
        a1[0] = 0; a1[1] = 1; a1[2] = 0; a1[3] = 0; a1[4] = dt / EST_MASS;
        zs[0] = p0 + dt * v0; zs[1] = v0 + dt / EST_MASS * zs[4] + dt * (-EST_GRAV - (k.fl + k.fr) / EST_MASS);
        par[16] = yL; par[17] = yR; par[18] = k.contact ? 1.0 : 0.0; par[19] = k.wm;   // for the terrain lag at the end
      }
    }
  ENDL
  // ---- covariance prediction P <- A P A' + Q (kalman_predict_cov): rows 0 and 1 of A P, column by column ...
  LANES
    const int f = l >> 3, j = l & 7, N = f < 2 ? 6 : 5;
    if (f < 3 && j < N) {
      double *P = f < 2 ? ds + ES_X + 42 * f + 6 : ds + ES_PZ; const double *a1 = ds + PAR + 64 * f;
      const double T0 = P[j] + EST_DT * P[N + j]; double T1 = 0;
      for (int q = 0; q < N; ++q) T1 += a1[q] * P[q * N + j];
      P[j] = T0; P[N + j] = T1;
    }
  ENDL
  LANES  // ... then (A P) A' row by row: columns 0 and 1 mix, the other diagonal entries receive Q
    const int f = l >> 3, i = l & 7, N = f < 2 ? 6 : 5;
    if (f < 3 && i < N) {
      double *P = f < 2 ? ds + ES_X + 42 * f + 6 : ds + ES_PZ; const double *a1 = ds + PAR + 64 * f, *Qd = a1 + 10; const double *T = P + i * N;
      double c1 = i == 1 ? Qd[1] : 0; for (int q = 0; q < N; ++q) c1 += T[q] * a1[q];
      const double c0 = ((i == 0 ? Qd[0] : 0) + T[0]) + T[1] * EST_DT;
      const double dg = i >= 2 ? Qd[i] + T[i] : 0;
      P[i * N] = c0; P[i * N + 1] = c1;
      if (i >= 2) P[i * N + i] = dg;
    }
  ENDL
  // ---- measurement update (kalman_update): rows are differences of two states or single states
  LANES  // H P (lane = column), P H' (lane = row)
    const int f = l >> 3, j = l & 7, N = f < 2 ? 6 : 5, K = f < 2 ? 4 : 2;
    if (f < 3 && j < N) {
      const double *P = f < 2 ? ds + ES_X + 42 * f + 6 : ds + ES_PZ; double *wk = ds + WRK + 112 * f;
      const int plus[4] = {0, 0, f < 2 ? 4 : 0, f < 2 ? 1 : 0}, minus[4] = {2, 3, -1, -1};
      for (int m = 0; m < K; ++m) {
        wk[W_HP + m * N + j] = minus[m] < 0 ? P[plus[m] * N + j] : P[plus[m] * N + j] - P[minus[m] * N + j];
        wk[W_G + j * K + m] = minus[m] < 0 ? P[j * N + plus[m]] : P[j * N + plus[m]] - P[j * N + minus[m]];
      }
    }
  ENDL
  LANES  // S = H (P H') + R, augmented with the identity (lane = column of [S | I])
    const int f = l >> 3, j = l & 7, K = f < 2 ? 4 : 2;
    if (f < 3 && j < 2 * K) {
      double *wk = ds + WRK + 112 * f;
      const int plus[4] = {0, 0, f < 2 ? 4 : 0, f < 2 ? 1 : 0}, minus[4] = {2, 3, -1, -1};
      const double Rd[4] = {1e-6, 1e-6, 1e-6, 1};
      for (int i = 0; i < K; ++i) {
        double a;
        if (j < K) { a = (i == j ? Rd[i] : 0) + wk[W_G + plus[i] * K + j]; if (minus[i] >= 0) a -= wk[W_G + minus[i] * K + j]; }
        else a = (i == j - K) ? 1.0 : 0.0;
        wk[W_S + i * 2 * K + j] = a;
      }
    }
  ENDL
  for (int c = 0; c < 4; ++c) {   // Gauss-Jordan without row exchanges (S is symmetric positive definite), one pivot per phase, lane = column
    LV(double, pv); LVA(double, col, 4);
    LANES
      const int f = l >> 3, j = l & 7, K = f < 2 ? 4 : 2;
      L(pv) = 0;
      if (f < 3 && j < 2 * K && c < K) {
        const double *S = ds + WRK + 112 * f + W_S;
        const double inv = 1.0 / S[c * 2 * K + c];
        L(pv) = S[c * 2 * K + j] * inv;
        for (int r = 0; r < K; ++r) LA(col, r) = r == c ? L(pv) : S[r * 2 * K + j] - S[r * 2 * K + c] * L(pv);
      }
    ENDL
    LANES
      const int f = l >> 3, j = l & 7, K = f < 2 ? 4 : 2;
      if (f < 3 && j < 2 * K && c < K) { double *S = ds + WRK + 112 * f + W_S; for (int r = 0; r < K; ++r) S[r * 2 * K + j] = LA(col, r); }
    ENDL
  }
  LANES  // gain K = (P H') inv(S) (lane = row)
    const int f = l >> 3, i = l & 7, N = f < 2 ? 6 : 5, K = f < 2 ? 4 : 2;
    if (f < 3 && i < N) {
      double *wk = ds + WRK + 112 * f;
      for (int j = 0; j < K; ++j) { double a = 0; for (int m = 0; m < K; ++m) a += wk[W_G + i * K + m] * wk[W_S + m * 2 * K + K + j]; wk[W_KG + i * K + j] = a; }
    }
  ENDL
  LV(double, xnew);
  LANES  // state: x += K (z - H x), the innovation taken from the predicted state before any entry is updated (lane = state entry)
    const int f = l >> 3, i = l & 7, N = f < 2 ? 6 : 5, K = f < 2 ? 4 : 2;
    L(xnew) = 0;
    if (f < 3 && i < N) {
      const double *wk = ds + WRK + 112 * f, *zm = ds + PAR + 64 * f + 6; const double *x = f < 2 ? ds + ES_X + 42 * f : ds + ES_Z;
      const int plus[4] = {0, 0, f < 2 ? 4 : 0, f < 2 ? 1 : 0}, minus[4] = {2, 3, -1, -1};
      double xi = x[i];
      for (int j = 0; j < K; ++j) { double inn = zm[j] - x[plus[j]]; if (minus[j] >= 0) inn += x[minus[j]]; xi += wk[W_KG + i * K + j] * inn; }
      L(xnew) = xi;
    }
  ENDL
  LANES  // covariance: P -= K (H P) (lane = column), and the new state
    const int f = l >> 3, j = l & 7, N = f < 2 ? 6 : 5, K = f < 2 ? 4 : 2;
    if (f < 3 && j < N) {
      double *P = f < 2 ? ds + ES_X + 42 * f + 6 : ds + ES_PZ; const double *wk = ds + WRK + 112 * f; double *x = f < 2 ? ds + ES_X + 42 * f : ds + ES_Z;
      for (int i = 0; i < N; ++i) { double a = 0; for (int m = 0; m < K; ++m) a += wk[W_KG + i * K + m] * wk[W_HP + m * N + j]; P[i * N + j] -= a; }
      x[j] = L(xnew);
    }
  ENDL
  // ---- terrain lag, outputs, filter row back to global memory
  LANES
    if (l == 16) { const double *par = ds + PAR + 128; const EstimatorContact k{par[18] != 0, 0, 0, par[19], 0, 0}; ds[ES_TERRAIN] = estimator_terrain(ds[ES_TERRAIN], ds[ES_Z], par[16], par[17], k); }
    if (l == 0) ds[0] = 1;
  ENDL
  LANES
    for (int i = l; i < ES_FORCE; i += 32) est[i] = ds[i];
    if (l < 2) { const double *xs = ds + ES_X + 42 * l; eo[EO_POS + l] = (real)xs[0]; eo[EO_VEL + l] = (real)xs[1]; eo[EO_EXTF + l] = (real)xs[5]; }
    if (l == 2) { const double *zs = ds + ES_Z; eo[EO_POS + 2] = (real)zs[0]; eo[EO_VEL + 2] = (real)zs[1]; eo[EO_EXTF + 2] = (real)zs[4]; eo[EO_TERRAIN] = (real)ds[ES_TERRAIN]; }
  ENDL
}

// ------------------------------------------------------------------ one control tick (cassie_sim_step_pd)
// soft-limit tables of the safety layer (cassie_core_sim_step), degrees; left leg then right leg
template <typename real> CFN real core_lo_deg(int i) { const real t[10] = {-15, -22, -50, -156, -140, -20, -22, -50, -156, -140}; return t[i]; }
template <typename real> CFN real core_hi_deg(int i) { const real t[10] = {20, 22, 80, -42, -35, 15, 22, 80, -42, -35}; return t[i]; }
template <typename real> CFN real core_K(int k) { const real t[5] = {1000, 800, 1200, 1200, 100}; return t[k]; }
template <typename real> CFN real core_C(int k) { const real t[5] = {12, 12, 36, 36, 7}; return t[k]; }

template <typename real, bool DR, int FEAT, bool EST = DR>
CFN void step_env(const DevModel<real> &cm, real *sm, const EnvPtrs<real> &E, LP(real, qvel), LP(real, qacc_ws), LP(real, xqvel), LP(real, xqacc_ws), int nticks, int mode) {
  const bool forward_only = (mode != 0);
  DECL_LANE
  real *cst = E.cst, *vecs = sm + S_VEC, *obs = E.obs; const real *pd = E.pd; int *ism = E.dfilt;
  LV(real, ctrl); LV(real, tq); LV(real, scale_part);
  // forward_only: a single mj_forward with zero ctrl (cassie_sim_init / reset, src/cassiemujoco.c:1029); it shares the one inlined copy of
  // mj_substep with the stepping path so the kernel's instruction footprint stays small
  if (forward_only) nticks = 1;
  const real *gait = E.gait; const int tick0 = ism[DF_TICK];
  for (int tick = 0; tick < nticks; ++tick) {
    if (forward_only) { LANES L(ctrl) = 0; ENDL }
    else {
    // ---- the controller stage works on a shared-memory copy of the environment's controller rows (sensor / filter state, FIR taps, PD row, observation row):
    // one burst of independent loads instead of a dozen dependent L2 round trips through the stage's phases.  The copy lives in the kinematics buffers,
    // which are dead between ticks; the estimator's filter data sits in the constraint-matrix region.
    real *cs = sm + S_XPOS, *ps = sm + S_XPOS + CST_W + DFILT_W, *os = sm + S_XPOS + CST_W + DFILT_W + 64; int *is = reinterpret_cast<int *>(sm + S_XPOS + CST_W);
    LANES
      real t[6], o[4] = {0, 0, 0, 0}; int u[3];
#pragma unroll
      for (int j = 0; j < 6; ++j) t[j] = cst[l + 32 * j];
#pragma unroll
      for (int j = 0; j < 3; ++j) u[j] = ism[l + 32 * j];
      const real p0 = pd[l], p1 = l + 32 < PD_W ? pd[l + 32] : real(0);
      if (obs) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (l + 32 * j < OBS_W) o[j] = obs[l + 32 * j];
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) cs[l + 32 * j] = t[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) is[l + 32 * j] = u[j];
      ps[l] = p0; if (l + 32 < 64) ps[l + 32] = p1;
#pragma unroll
      for (int j = 0; j < 4; ++j) if (l + 32 * j < OBS_W) os[l + 32 * j] = o[j];
    ENDL
    // ---- pd_input_step (motor-PD branch) + cassie_core_sim_step, lane = motor; both read LAST tick's cassie_out
    const real W = real(0.15), DEG = real(0.017453292519943295);
    LANES
      L(tq) = 0; L(scale_part) = 1;
      if (l < 10) {
        const real pos = cs[CS_DPOS + l], vel = cs[CS_DVEL + l]; const int k = l % 5;
        real pT = ps[10 + l];
        if (gait) pT += gait[GA_AMP + l] * msin(real(6.283185307179586) * gait[GA_FREQ] * ((real)(tick0 + tick) * real(0.0005)) + gait[GA_PHASE + l]);   // open-loop gait (BASELINE config 5)
        const real u = ps[l] + ps[30 + l] * (pT - pos) + ps[40 + l] * (ps[20 + l] - vel);
        real add = 0, sc = 1;
        const real dhi = pos - (core_hi_deg<real>(l) * DEG - W), dlo = (core_lo_deg<real>(l) * DEG + W) - pos;
        if (dhi > 0) { add -= core_K<real>(k) * dhi * (1 + dhi / W) + core_C<real>(k) * mmin(dhi / W, real(1)) * vel; sc *= mmax(real(0), 1 - dhi / W); }
        if (dlo > 0) { add += core_K<real>(k) * dlo * (1 + dlo / W) - core_C<real>(k) * mmin(dlo / W, real(1)) * vel; sc *= mmax(real(0), 1 - dlo / W); }
        if (k == 2 || k == 3) {  // coupled row hipPitch + knee >= -135 deg, evaluated by both lanes of the pair
          const int a = l - k + 2; const real dsum = -135 * DEG - (cs[CS_DPOS + a] + cs[CS_DPOS + a + 1]);
          if (dsum > 0) { add += 1200 * dsum * (1 + dsum / W) - 36 * mmin(dsum / W, real(1)) * vel; if (k == 2) sc *= mmax(real(0), 1 - dsum / W); }
        }
        L(tq) = u; L(ctrl) = add; L(scale_part) = sc;
      }
    ENDL
    // ---- pd_input_step, taskPd branch (closed, decoded): task-space PD on the foot point through the leg Jacobian; extended instance only
    if (DR && E.task) {
      const real *tk = E.task;
      LANES  // lanes 0..13: sin / cos of the 14 chain angles from LAST tick's cassie_out
        if (l < 14) {
          const int i = l, sd = i / 7, k = i % 7; const real *mp = cs + CS_DPOS + 5 * sd, *jp = cs + CS_JPOS + 3 * sd;
          real a = k < 3 ? mp[k] : mp[3];
          if (k >= 4) a += jp[0]; if (k >= 5) a += jp[1]; if (k >= 6) a += mp[4];
          real sn, cs; msincos(a, &sn, &cs); vecs[32 + 2 * i] = sn; vecs[33 + 2 * i] = cs;
        }
      ENDL
      LANES
        if (l < 2) {
          const int sd = l; const real *mv = cs + CS_DVEL + 5 * sd, *jv = cs + CS_JVEL + 3 * sd, *t = tk + 30 * sd;
          const real rate[7] = {mv[0], mv[1], mv[2], mv[3], jv[0], jv[1], mv[4]};
          real fo[13], jac[30], x[6], w[6];
          est_foot<real>(sd, vecs + 32 + 14 * sd, rate, fo, jac);
          const real qw = fo[3], qx = fo[4], qy = fo[5], qz = fo[6];
          x[0] = fo[0]; x[1] = fo[1]; x[2] = fo[2];
          x[3] = matan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz)); x[4] = masin(2 * (qw * qy - qz * qx)); x[5] = matan2(2 * (qw * qx + qy * qz), 1 - 2 * (qx * qx + qy * qy));
          for (int k = 0; k < 6; ++k) { const real vk = k < 3 ? fo[10 + k] : fo[7 + k - 3]; w[k] = t[k] + t[18 + k] * (t[6 + k] - x[k]) + t[24 + k] * (t[12 + k] - vk); }
          for (int j = 0; j < 5; ++j) { real acc = 0; for (int k = 0; k < 6; ++k) acc += jac[6 * j + k] * w[k]; vecs[64 + 5 * sd + j] = acc; }
        }
      ENDL
      LANES if (l < 10) L(tq) += vecs[64 + l]; ENDL
    }
    // product of the per-row scale factors over all 10 motors (log-free: multiply through shared memory)
    LANES vecs[l] = L(scale_part); ENDL
    LANES
      if (l < 10) {
        real sc = 1; for (int i = 0; i < 10; ++i) sc *= vecs[i];
        const bool sto = !(cs[CS_STO] >= 1);
        real t = sto ? real(0) : L(tq) * sc + L(ctrl);
        const real lim = cm.act_torque_limit[l];
        t = clampr(t, -lim, lim);
        // ---- motor(): torque-speed curve, STO, 6-tick delay line (src/cassiemujoco.c:638-664)
        const real ratio = cm.act_gear[l], tmax = cm.act_ctrl_hi[l], w = cs[CS_ACTVEL + l];
        real tlim = 2 * tmax * (1 - mabs(w) / cm.act_wmax[l]); tlim = mmax(mmin(tlim, tmax), real(0));
        if (sto) t = 0;
        real tau = mmin(mabs(t / ratio), tlim); if (t < 0) tau = -tau;
        real *dl = cs + CS_DELAY + 6 * l;
        const real c = dl[5];
        for (int k = 5; k > 0; --k) dl[k] = dl[k - 1];
        dl[0] = tau;
        L(ctrl) = c;
        cs[CS_DTORQUE + l] = c * ratio;
      } else L(ctrl) = 0;
    ENDL
    // ---- cassie_sensor_data(): encoders + filters from the sensordata of the PREVIOUS physics step (:737-774)
    LANES
      if (l < 10) {  // drive encoder, integer FIR (:558-593)
        const int s = l < 5 ? l : l + 3, bits = cm.enc_bits[s]; int *x = is + 9 * l;
        const double TWO_PI = 6.283185307179586;
        const int enc = (int)((double)cs[CS_SENSOR + s] / TWO_PI * (double)(1 << bits));
        const double scale = TWO_PI / (double)(1 << bits) / (double)cm.act_gear[l];
        cs[CS_DPOS + l] = (real)(enc * scale);
        bool allzero = true; for (int k = 0; k < 9; ++k) allzero &= (x[k] == 0);
        if (allzero) for (int k = 0; k < 9; ++k) x[k] = enc;
        for (int k = 8; k > 0; --k) x[k] = x[k - 1];
        x[0] = enc;
        const int b[9] = {2727, 534, -2658, -795, 72, 110, 19, -6, -3};
        int y = 0; for (int k = 0; k < 9; ++k) y += x[k] * b[k];
        cs[CS_DVEL + l] = (real)(y * scale / 3.141592653589793);
      } else if (l < 16) {  // joint encoder, IIR (:596-635)
        const int i = l - 10, s = i < 3 ? 5 + i : 10 + i, bits = cm.enc_bits[s]; real *x = cs + (CS_JFX - 40) + 4 * l, *y = cs + (CS_JFY - 30) + 3 * l;   // = CS_JFX + 4 i, CS_JFY + 3 i, written with non-negative terms only
        const double TWO_PI = 6.283185307179586;
        const int enc = (int)((double)cs[CS_SENSOR + s] / TWO_PI * (double)(1 << bits));
        const real p = (real)(enc * (TWO_PI / (double)(1 << bits)));
        cs[CS_JPOS + i] = p;
        bool allzero = true; for (int k = 0; k < 4; ++k) allzero &= (x[k] == 0);
        if (allzero) for (int k = 0; k < 4; ++k) x[k] = p;
        for (int k = 3; k > 0; --k) x[k] = x[k - 1];
        x[0] = p;
        for (int k = 2; k > 0; --k) y[k] = y[k - 1];
        real y0 = real(12.348) * (x[0] + x[1] - x[2] - x[3]);
        y0 -= y[1] * real(-1.7658) + y[2] * real(0.79045);
        y[0] = y0; cs[CS_JVEL + i] = y0;
      }
    ENDL
    // ---- *y = cassie_out (:1127): the observation of this tick
    bool est_on = false;
    if constexpr (EST) est_on = E.est != (double *)0;   // the filters advance every 2 kHz tick, so the stateless part runs every tick too
    const bool wrote_obs = obs && (tick == nticks - 1 || est_on);
    if (wrote_obs) {
      LANES
        if (l < 10) { os[OB_MPOS + l] = cs[CS_DPOS + l]; os[OB_MVEL + l] = cs[CS_DVEL + l]; os[OB_MTORQUE + l] = cs[CS_DTORQUE + l]; }
        if (l < 6) { os[OB_JPOS + l] = cs[CS_JPOS + l]; os[OB_JVEL + l] = cs[CS_JVEL + l]; }
        if (l < 13) os[OB_QUAT + l] = cs[CS_SENSOR + 16 + l];
        if (l == 13) os[OB_TIME] = cs[CS_TIME];
      ENDL
      // ---- state_output_step, stateless part (closed source, decoded): pelvis orientation / acceleration, foot poses and velocities
      LANES  // lanes 0..13: sin / cos of the 14 chain angles (3 hip angles + 4 cumulative planar angles per leg)
        if (l < 14) {
          const int i = l, sd = i / 7, k = i % 7; const real *mp = cs + CS_DPOS + 5 * sd, *jp = cs + CS_JPOS + 3 * sd;
          real a = k < 3 ? mp[k] : mp[3];
          if (k >= 4) a += jp[0]; if (k >= 5) a += jp[1]; if (k >= 6) a += mp[4];
          real sn, cs; msincos(a, &sn, &cs); vecs[32 + 2 * i] = sn; vecs[33 + 2 * i] = cs;
        }
      ENDL
      LANES
        if (l < 2) {
          const int sd = l; const real *mv = cs + CS_DVEL + 5 * sd, *jv = cs + CS_JVEL + 3 * sd;
          const real rate[7] = {mv[0], mv[1], mv[2], mv[3], jv[0], jv[1], mv[4]};
          est_foot<real>(sd, vecs + 32 + 14 * sd, rate, os + OB_FOOT + 13 * sd);
        } else if (l == 2) {
          const real *q = cs + CS_SENSOR + 16, *w = cs + CS_SENSOR + 20, *a = cs + CS_SENSOR + 23;
          real R[9], wr[3], wwr[3]; const real r[3] = {real(0.03155), 0, real(-0.079996)};
          quat2mat(R, q); cross3(wr, w, r); cross3(wwr, w, wr);
          est_mat2quat(os + OB_EST_QUAT, R);   // pelvis.orientation: the IMU quaternion through its matrix and back (+-q, mat2quat's sign)
          for (int k = 0; k < 3; ++k) os[OB_EST_ACC + k] = a[k] - R[6 + k] * real(9.806) - wwr[k];
        }
      ENDL
      // ---- state_output_step, spring-force model and filters: the warp-parallel stage above (extended instance only)
      if constexpr (EST) { if (est_on) est_stage<real>(sm, cs, os, E.est, vecs + 32); }
    }
    // ---- flush: the rows' modified ranges go back to their HBM rows before the physics (which writes sensordata / actuator_velocity / time there)
    LANES
      for (int i = CS_DPOS + l; i < CS_TIME; i += 32) cst[i] = cs[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) { const int i = l + 32 * j; if (i < DF_TICK) ism[i] = is[i]; }
      if (wrote_obs) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const int i = l + 32 * j; if (i < OBS_W) obs[i] = os[i]; }
      }
    ENDL
    }
    // ---- mj_step1 + mj_step2, round(5e-4 / timestep) times with ctrl held (:1130-1134)
    const int nsub = forward_only ? 1 : (E.nsub > 0 ? E.nsub : cm.nsub);
    for (int s = 0; s < nsub; ++s) mj_substep<real, DR, FEAT>(cm, sm, E, qvel, qacc_ws, xqvel, xqacc_ws, ctrl, (tick == nticks - 1 && s == nsub - 1) ? E.dbg : (real *)0, (tick == nticks - 1 && s == nsub - 1) ? E.aux : (real *)0, mode);
  }
  if (!forward_only) { LANES if (l == 0) ism[DF_TICK] = tick0 + nticks; ENDL }
}

}  // namespace cassie
