// cassie_b200.cu -- CUDA kernels (sm_100a) and the C-ABI of the batched Cassie stepper.  See include/cassie_b200.h.
//
// Kernel shape: one warp per environment; a CTA of WPB warps first stages the model constant block into shared memory with
// a TMA bulk copy (cp.async.bulk + mbarrier, SASS: UBLKCP), then every warp loads its environment's state rows from HBM
// (coalesced, one row per array), advances `nticks` control ticks entirely on chip (step_core.inl) and writes the rows back.
#include <cuda_runtime.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <omp.h>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/cassie_b200.h"
#include "devbuild.h"
#include "estimator_host.h"
#include "step_kernel.cuh"

// the step-kernel instances live in their own translation units (step_inst.cu, one compilation per instance)
#define CASSIE_STEP_ENTRY(tag) extern "C" const void *cassie_step_entry_##tag(void);
CASSIE_STEP_ENTRY(f00) CASSIE_STEP_ENTRY(f10) CASSIE_STEP_ENTRY(f02) CASSIE_STEP_ENTRY(f12) CASSIE_STEP_ENTRY(f04) CASSIE_STEP_ENTRY(f14) CASSIE_STEP_ENTRY(f05) CASSIE_STEP_ENTRY(f15) CASSIE_STEP_ENTRY(f07) CASSIE_STEP_ENTRY(f17)
CASSIE_STEP_ENTRY(d07) CASSIE_STEP_ENTRY(d17)
CASSIE_STEP_ENTRY(f20) CASSIE_STEP_ENTRY(f22) CASSIE_STEP_ENTRY(f24) CASSIE_STEP_ENTRY(f25) CASSIE_STEP_ENTRY(f27) CASSIE_STEP_ENTRY(d27)
#undef CASSIE_STEP_ENTRY

namespace cassie {

// the instance for (precision, plain / extended, model features): the smallest compiled feature set that covers the model's
template <typename real> static const void *step_entry(int inst, int feat, int *compiled_feat = nullptr) {   // inst: 0 plain, 1 extended, 2 plain + estimator
  int f = F_ALL;
  if (!std::is_same<real, double>::value) { if (feat == 0) f = 0; else if (feat == F_HFIELD) f = F_HFIELD; else if (feat == F_BOX) f = F_BOX; else if ((feat & ~(F_XB | F_BOX)) == 0) f = F_XB | F_BOX; }
  if (compiled_feat) *compiled_feat = f;
  if (std::is_same<real, double>::value) return inst == 1 ? cassie_step_entry_d17() : (inst == 2 ? cassie_step_entry_d27() : cassie_step_entry_d07());
  switch (f) {
    case 0: return inst == 1 ? cassie_step_entry_f10() : (inst == 2 ? cassie_step_entry_f20() : cassie_step_entry_f00());
    case F_HFIELD: return inst == 1 ? cassie_step_entry_f12() : (inst == 2 ? cassie_step_entry_f22() : cassie_step_entry_f02());
    case F_BOX: return inst == 1 ? cassie_step_entry_f14() : (inst == 2 ? cassie_step_entry_f24() : cassie_step_entry_f04());
    case F_XB | F_BOX: return inst == 1 ? cassie_step_entry_f15() : (inst == 2 ? cassie_step_entry_f25() : cassie_step_entry_f05());
    default: return inst == 1 ? cassie_step_entry_f17() : (inst == 2 ? cassie_step_entry_f27() : cassie_step_entry_f07());
  }
}

static thread_local std::string g_err;
static void set_err(const std::string &e) { g_err = e; fprintf(stderr, "cassie_b200: %s\n", e.c_str()); }
#define CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(std::string(#call) + ": " + cudaGetErrorString(e_)); return false; } } while (0)

// CPUs this process may actually use: the affinity mask clipped by the cgroup CPU quota (a container on a 128-thread host may own 16)
static int effective_cpus() {
  int hw = omp_get_num_procs();
  long quota = -1, period = 0;
  if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) { char q[64]; if (fscanf(f, "%63s %ld", q, &period) == 2 && strcmp(q, "max")) quota = atol(q); fclose(f); }
  else {
    if (FILE *g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) { if (fscanf(g, "%ld", &quota) != 1) quota = -1; fclose(g); }
    if (FILE *g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (fscanf(g, "%ld", &period) != 1) period = 0; fclose(g); }
  }
  if (quota > 0 && period > 0) { const int c = (int)((quota + period - 1) / period); if (c >= 1 && c < hw) hw = c; }
  return hw < 1 ? 1 : hw;
}
// host pack / unpack threads of the AoS entry point.  The OpenMP workers busy-wait between the call's parallel regions, so the team burns its CPUs for the
// whole call: by default two of the CPUs the process may use (affinity mask clipped by the cgroup quota) are left to the CUDA driver's and the caller's
// own threads -- a team as large as the quota gets the whole cgroup throttled (measured on an 8-GPU lease: 8 x 12 threads on 96 CPUs, 4 x slower loops)
static int aos_thread_count() {
  const char *e = getenv("CASSIE_B200_AOS_THREADS"); const int hw = effective_cpus();
  int t = e ? atoi(e) : (hw > 4 ? hw - 2 : hw); if (t > 32) t = 32; if (t > hw) t = hw; return t < 1 ? 1 : t;
}

// cassie_integrate_pos for the whole batch (mj_integratePos, src/cassiemujoco.c:1183-1189): the HBM-bound kernel.
// A tile of ITILE consecutive environments is one contiguous chunk of the qpos array and one of the qvel array.  Persistent CTAs
// stream tiles through an ISTAGES-deep shared-memory ring: TMA bulk loads (cp.async.bulk -> mbarrier), in-place arithmetic,
// TMA bulk store of the qpos chunk.  Algorithmic traffic per env: read qpos + qvel, write qpos = (36 + 32 + 36) reals.
constexpr int ITILE = 64, ISTAGES = 4;
template <typename real> __host__ __device__ constexpr size_t itile_bytes(int qw, int vw) { return (size_t)ITILE * (qw + vw) * sizeof(real); }
template <typename real>
__global__ void __launch_bounds__(256) cassie_integrate_kernel(const DevModel<real> *__restrict__ gmodel, real *__restrict__ qpos, const real *__restrict__ qvel, int n, int QPOS_W, int QVEL_W) {
  extern __shared__ __align__(128) unsigned char ibuf[];
  __shared__ __align__(8) uint64_t bar[ISTAGES];
  __shared__ int ns, nb, sq_adr[MJ + 4], sd_adr[MJ + 4], bq_adr[MJ], bd_adr[MJ];
  __shared__ real h;
  const int ntiles = (n + ITILE - 1) / ITILE;
  if (threadIdx.x == 0) {
    int a = 0, b = 0;
    for (int j = 0; j < gmodel->njnt; ++j) {
      const int t = gmodel->jnt_type[j];
      if (t >= 2) { sq_adr[a] = gmodel->jnt_qposadr[j]; sd_adr[a] = gmodel->jnt_dofadr[j]; ++a; } else if (t == 1) { bq_adr[b] = gmodel->jnt_qposadr[j]; bd_adr[b] = gmodel->jnt_dofadr[j]; ++b; }
      else {   // free joint: three translations like slides, the quaternion like a ball joint
        for (int k = 0; k < 3; ++k) { sq_adr[a] = gmodel->jnt_qposadr[j] + k; sd_adr[a] = gmodel->jnt_dofadr[j] + k; ++a; }
        bq_adr[b] = gmodel->jnt_qposadr[j] + 3; bd_adr[b] = gmodel->jnt_dofadr[j] + 3; ++b;
      }
    }
    ns = a; nb = b; h = gmodel->timestep;
    for (int s = 0; s < ISTAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue_load = [&](int stage, int tile) {   // thread 0 only
    const int cnt = min(ITILE, n - tile * ITILE);
    const uint32_t bq = (uint32_t)(cnt * QPOS_W * sizeof(real)), bv = (uint32_t)(cnt * QVEL_W * sizeof(real));
    unsigned char *dst = ibuf + (size_t)stage * itile_bytes<real>(QPOS_W, QVEL_W);
    const uint32_t bar_a = smem_u32(&bar[stage]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bq + bv) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(qpos + (size_t)tile * ITILE * QPOS_W), "r"(bq), "r"(bar_a) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + (size_t)ITILE * QPOS_W * sizeof(real))), "l"(qvel + (size_t)tile * ITILE * QVEL_W), "r"(bv), "r"(bar_a) : "memory");
  };
  if (threadIdx.x == 0) for (int s = 0; s < ISTAGES; ++s) { const int tile = blockIdx.x + s * gridDim.x; if (tile < ntiles) issue_load(s, tile); }
  for (int i = 0;; ++i) {
    const int tile = blockIdx.x + i * gridDim.x;
    if (tile >= ntiles) break;
    const int stage = i % ISTAGES; const uint32_t parity = (uint32_t)((i / ISTAGES) & 1), bar_a = smem_u32(&bar[stage]);
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(bar_a), "r"(parity) : "memory");
    real *sq = reinterpret_cast<real *>(ibuf + (size_t)stage * itile_bytes<real>(QPOS_W, QVEL_W)); const real *sv = sq + ITILE * QPOS_W;
    const int cnt = min(ITILE, n - tile * ITILE);
    // scalar joints: lane = joint, warps stride over the tile's environments (a row's 36 entries sit in distinct banks);
    // ball joints: thread = (env, ball joint), dense, no divergence
    {
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
      if (lane < ns) { const int qa = sq_adr[lane], da = sd_adr[lane]; for (int e = wid; e < cnt; e += nw) sq[e * QPOS_W + qa] += h * sv[e * QVEL_W + da]; }
      for (int w = threadIdx.x; w < cnt * nb; w += blockDim.x) {
        const int e = w / nb, k = w - e * nb;
        real *q = sq + e * QPOS_W + bq_adr[k]; const real *v = sv + e * QVEL_W + bd_adr[k];
        real wv[3] = {v[0], v[1], v[2]}, qq[4] = {q[0], q[1], q[2], q[3]}, qr[4], s, c;
        const real ang = h * normalize3(wv);
        msincos(real(0.5) * ang, &s, &c); qr[0] = c; qr[1] = wv[0] * s; qr[2] = wv[1] * s; qr[3] = wv[2] * s;
        if (ang == 0) { qr[0] = 1; qr[1] = qr[2] = qr[3] = 0; }
        normalize4(qq); mul_quat(qq, qq, qr);
        q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk store
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(qpos + (size_t)tile * ITILE * QPOS_W), "r"(smem_u32(sq)), "r"((uint32_t)(cnt * QPOS_W * sizeof(real))) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      const int next = tile + ISTAGES * gridDim.x;
      if (next < ntiles) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); issue_load(stage, next); }
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// rows of the selected environments <- one template row (masked reset / set_const)
template <typename T>
__global__ void fill_rows_kernel(T *__restrict__ dst, const T *__restrict__ row, int w, const unsigned char *__restrict__ mask, int n, int ncols) {
  const size_t total = (size_t)n * w;   // only the first ncols columns of a row are rewritten
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) { const size_t e = i / w; if ((!mask || mask[e]) && (int)(i - e * w) < ncols) dst[i] = row[i - e * w]; }
}

// rows of the selected environments <- the same rows of another array (snapshot restore of a subset)
template <typename T>
__global__ void copy_rows_kernel(T *__restrict__ dst, const T *__restrict__ src, int w, const unsigned char *__restrict__ mask, int n) {
  const size_t total = (size_t)n * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) { const size_t e = i / w; if (!mask || mask[e]) dst[i] = src[i]; }
}

// ------------------------------------------------------------------ host side
// device-resident copy of every per-environment row array of a batch (cassie_state_t / cassie_batch_get_state); it owns its memory and carries its own
// sizes, so it can be copied, restored into any batch of the same shape and freed without the batch it was taken from
struct Snap {
  enum { R_QPOS, R_QVEL, R_QACC, R_CST, R_XFRC, R_OBS, R_DFILT, R_EST, R_AUX, NROWS };
  int device = 0, n = 0, esz = 4, qw = 0, vw = 0; bool has_est = false, has_aux = false;
  void *rows[NROWS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int width(int r) const { const int w[NROWS] = {qw, vw, vw, CST_W, XFRC_W, OBS_W, DFILT_W, EST_W, AUX_W}; return w[r]; }
  int elem(int r) const { return r == R_DFILT ? 4 : (r == R_EST ? 8 : esz); }
  size_t bytes(int r) const { return (size_t)n * width(r) * elem(r); }
};
static void snap_delete(Snap *k) { if (!k) return; cudaSetDevice(k->device); for (int r = 0; r < Snap::NROWS; r++) cudaFree(k->rows[r]); delete k; }
static Snap *snap_new(int device, int n, int esz, int qw, int vw) {
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  Snap *k = new Snap(); k->device = device; k->n = n; k->esz = esz; k->qw = qw; k->vw = vw;
  for (int r = 0; r < Snap::NROWS; r++) if (r != Snap::R_EST && r != Snap::R_AUX && cudaMalloc(&k->rows[r], k->bytes(r)) != cudaSuccess) { snap_delete(k); return nullptr; }
  return k;
}
static bool snap_clone_into(Snap *d, const Snap *k) {   // same shape required
  if (d->n != k->n || d->esz != k->esz || d->qw != k->qw || d->vw != k->vw || cudaSetDevice(d->device) != cudaSuccess) return false;
  d->has_est = k->has_est; d->has_aux = k->has_aux;
  for (int r = 0; r < Snap::NROWS; r++) {
    if ((r == Snap::R_EST && !k->has_est) || (r == Snap::R_AUX && !k->has_aux)) continue;
    if (!d->rows[r] && cudaMalloc(&d->rows[r], d->bytes(r)) != cudaSuccess) return false;
    if (cudaMemcpy(d->rows[r], k->rows[r], k->bytes(r), cudaMemcpyDefault) != cudaSuccess) return false;
  }
  return true;
}

struct BatchBase {
  virtual ~BatchBase() {}
  bool est_forces = false;   // fill toeForce / heelForce of state_out_t on the host (cassie_batch_enable_estimator_forces)
  bool est_filter = false;   // run the estimator's filters on the host, one per environment (cassie_batch_enable_estimator_filter; needs est_forces): the checker of the in-kernel estimator
  double aos_t[6] = {0, 0, 0, 0, 0, 0}; cudaEvent_t aos_ev[3] = {nullptr, nullptr, nullptr};   // [4] [5]: H2D and kernel milliseconds by CUDA events, only with CASSIE_B200_AOS_EVENTS set
  double aos_unused[1] = {0};   // accumulated seconds of the AoS entry point: host pack, device (copies + kernel, as waited for), host unpack; [3] calls
  bool est_auto = true;      // cassie_sim_step_pd_batch switches the in-kernel estimator on by itself (cassie_batch_enable_estimator_device(b, 0) clears this)
  std::vector<cassie::EstimatorFilter> est_state;
  void reset_estimator(const unsigned char *mask) { for (size_t e = 0; e < est_state.size(); e++) if (!mask || mask[e]) est_state[e].reset(); }
  HostModel hm; int n = 0, device = 0, precision = 0, wpb = 4; cudaStream_t stream = nullptr; bool own_stream = false; long launches = 0; bool debug = false;
  virtual bool init() = 0;
  virtual bool reset(const unsigned char *mask) = 0;
  virtual bool set_pd(const double *pd) = 0;
  virtual bool step(int nticks, int mode) = 0;
  virtual bool integrate() = 0;
  virtual bool get(const char *field, double *out) = 0;   // qpos[n][35] qvel[n][32] time[n] obs[n][64]
  virtual bool set(const char *field, const double *in) = 0;
  virtual bool set_xfrc(const double *xfrc, int body) = 0;
  virtual void *dev_ptr(const char *field) = 0;
  virtual bool get_counters(int *out) = 0;
  virtual int debug_dump(int env, double *out, int cnt) = 0;
  virtual bool step_pd_aos(const pd_in_t *pd_in, state_out_t *state_out, const double *radio) = 0;
  virtual bool set_hfield(const float *data, int n_terrains) = 0;
  virtual bool enable_aux(bool on) = 0;
  virtual bool enable_estimator_device(bool on) = 0;   // leg forces + filters inside the step kernel (extended instance), every 2 kHz tick
  virtual bool reset_estimator_device(const unsigned char *mask) = 0;
  virtual bool set_task_pd(const double *rows) = 0;   // [n][60] or null (off)
  virtual bool set_pd_gait(const double *amp, const double *freq, const double *phase) = 0;   // [n][10], [n], [n][10] or all null (off)
  virtual bool set_model_rows(const char *what, const double *rows, int width) = 0;   // per-env model constants (domain randomisation)
  virtual bool get_model_rows(const char *what, double *rows, int width) = 0;
  virtual bool set_const(const unsigned char *mask, bool reset_state) = 0;
  virtual bool has_aux() const = 0;
  // full dynamic state of every environment (cassie_get_state / cassie_set_state, src/cassiemujoco.c:3435-3452: mjData + the block objects +
  // cassie_out + encoder filters + torque delay line): device-resident snapshots, restore of a masked subset
  virtual Snap *snap_alloc() = 0;
  virtual bool snap_get(Snap *snap) = 0;
  virtual bool snap_set(const Snap *snap, const unsigned char *mask) = 0;
  virtual bool copy_model_from(BatchBase *src) = 0;     // per-environment constants, height fields, timestep (mj_copyModel in cassie_sim_copy, :1083-1091)
  virtual bool rebuild_model() = 0;                     // after hm was edited (timestep, hold / release): rebuild and upload the constant block
  int nsub_override = 0;                                // > 0: physics sub-steps per control tick of the next launches (cassie_sim_step_pd_no2khz)
  bool sync() { CUDA_OK(cudaStreamSynchronize(stream)); return true; }
};

template <typename real> struct Batch : BatchBase {
  DevModel<real> *d_model = nullptr; EnvArrays<real> A{}; int QW = QPOS_W_MAIN, VW = QVEL_W_MAIN;
  struct LaunchCfg { int wpb = 1; size_t smem = 0; int resident_ctas = 1; } cfg[3];   // [0] plain instance, [1] extended instance, [2] plain + estimator
  int feat = F_ALL;   // model features (F_XB | F_HFIELD | F_BOX): selects the kernel instance
  std::vector<real> h_tmp;
  real *pin_pd = nullptr, *pin_obs = nullptr, *pin_task = nullptr;   // pinned staging for the AoS entry point
  real *dpin_pd = nullptr, *dpin_obs = nullptr;                      // the same buffers as the device sees them (mapped)
  static constexpr int AOS_MAXC = 4;
  cudaEvent_t aos_done[AOS_MAXC] = {}, aos_start = nullptr;            // the AoS entry point's chunk launches: completion events,
  cudaStream_t aos_stream[AOS_MAXC] = {};                              // their streams
  bool task_from_aos = false;   // the task-PD rows were created by the AoS entry point's forwarding (not by cassie_batch_set_task_pd)
  float *d_hfield = nullptr; unsigned char *d_mask = nullptr; DevModel<real> h_model_copy{}; int geom_dev[256]; void *d_row = nullptr;
  ~Batch() override {
    cudaSetDevice(device);
    cudaFree(d_model); cudaFree(A.qpos); cudaFree(A.qvel); cudaFree(A.qacc_ws); cudaFree(A.cst); cudaFree(A.pd); cudaFree(A.xfrc); cudaFree(A.obs); cudaFree(A.dbg);
    cudaFree(A.dfilt); cudaFree(A.counters); cudaFree(A.qM); cudaFree(A.ticket); cudaFree(A.aux); cudaFree(A.cenv); cudaFree(A.task); cudaFree(A.gait); cudaFree(A.est); if (pin_task) cudaFreeHost(pin_task); cudaFree(d_mask); cudaFree(d_row);
    for (int i = 0; i < AOS_MAXC; i++) { if (aos_done[i]) cudaEventDestroy(aos_done[i]); if (aos_stream[i]) cudaStreamDestroy(aos_stream[i]); } if (aos_start) cudaEventDestroy(aos_start);
    if (pin_pd) cudaFreeHost(pin_pd); if (pin_obs) cudaFreeHost(pin_obs); if (d_hfield) cudaFree(d_hfield);
    if (own_stream && stream) cudaStreamDestroy(stream);
  }
  bool init() override {
    CUDA_OK(cudaSetDevice(device));
    DevModel<real> *hmodel = (DevModel<real> *)calloc(1, model_bytes<real>()); std::string err; BuildInfo info;
    if (!build_dev_model(hm, *hmodel, err, &info)) { free(hmodel); set_err(err); return false; }
    static bool noted = false;   // once per process
    if (info.unsupported_pairs && !noted) noted = true, fprintf(stderr, "cassie_b200: note: %d candidate geom pairs involve box/hfield geoms that this build does not collide (skipped)\n", info.unsupported_pairs);
    CUDA_OK(cudaMalloc(&d_model, model_bytes<real>()));
    CUDA_OK(cudaMemcpy(d_model, hmodel, model_bytes<real>(), cudaMemcpyHostToDevice));
    h_model_copy = *hmodel; memcpy(geom_dev, info.geom_dev, sizeof geom_dev);
    feat = hmodel->xb >= 0 ? F_XB : 0;   // which kernel instance this model needs (step_inst.cu)
    for (int p = 0; p < hmodel->npair; p++) { const int k = pair_kind(hmodel->pair_code[p]); if (k == PAIR_HFIELD_SPHERE || k == PAIR_HFIELD_CAPSULE) feat |= F_HFIELD; else if (k >= PAIR_PLANE_BOX) feat |= F_BOX; }
    const int hmodel_ystride = hmodel->ystride; QW = hmodel->qpos_w; VW = hmodel->qvel_w; A.qpos_w = QW; A.qvel_w = VW; A.ystride = hmodel_ystride; A.xb = hmodel->xb; free(hmodel);
    A.n = n; A.hfield = nullptr; A.n_terrain = 1; A.hfield_stride = 0; { const char *e = getenv("CASSIE_B200_SYNCMASK"); A.cta_sync = getenv("CASSIE_B200_NOSYNC") ? 0 : (e ? (atoi(e) & 127) : 10); }   // which of the five per-sub-step rendezvous are on
    if (hm.nhfield == 1 && !set_hfield(nullptr, 1)) return false;
    CUDA_OK(cudaMalloc(&A.qpos, sizeof(real) * n * QW)); CUDA_OK(cudaMalloc(&A.qvel, sizeof(real) * n * VW)); CUDA_OK(cudaMalloc(&A.qacc_ws, sizeof(real) * n * VW));
    CUDA_OK(cudaMalloc(&A.cst, sizeof(real) * n * CST_W)); CUDA_OK(cudaMalloc(&A.pd, sizeof(real) * n * PD_W)); CUDA_OK(cudaMalloc(&A.xfrc, sizeof(real) * n * XFRC_W));
    CUDA_OK(cudaMalloc(&A.obs, sizeof(real) * n * OBS_W)); CUDA_OK(cudaMalloc(&A.dfilt, sizeof(int) * n * DFILT_W)); CUDA_OK(cudaMalloc(&A.counters, sizeof(int) * n * 8)); CUDA_OK(cudaMalloc(&A.qM, sizeof(real) * n * 2 * NM_MAX)); CUDA_OK(cudaMalloc(&A.ticket, sizeof(int)));
    CUDA_OK(cudaMemset(A.pd, 0, sizeof(real) * n * PD_W)); CUDA_OK(cudaMemset(A.obs, 0, sizeof(real) * n * OBS_W)); CUDA_OK(cudaMemset(A.counters, 0, sizeof(int) * n * 8));
    if (debug) { CUDA_OK(cudaMalloc(&A.dbg, sizeof(real) * n * D_SIZE)); CUDA_OK(cudaMemset(A.dbg, 0, sizeof(real) * n * D_SIZE)); }
    CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)); own_stream = true;
    // warps per CTA: maximise resident warps per SM over 1..4 CTAs per SM (each CTA carries its own copy of the model block); the two kernel
    // instances have different scratch sizes, hence their own launch shapes
    int dev_smem = 0, sm_smem = 0, sms = 0;
    CUDA_OK(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CUDA_OK(cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    for (int ext = 0; ext < 3; ++ext) {
      const long wb = (long)warp_bytes<real>(hmodel_ystride, ext == 1), mb = (long)model_bytes<real>() + 256;   // + static shared (mbarrier) and alignment slack
      const char *w = getenv("CASSIE_B200_WPB"); int k_sel = 1;
      if (w) { k_sel = atoi(w); const int kmax = (int)(((long)dev_smem - mb) / wb); if (k_sel > kmax) k_sel = kmax; }
      else {
        int best = 0, kk[5] = {0, 0, 0, 0, 0};
        for (int ctas = 4; ctas >= 1; --ctas) {
          long per_cta = (long)sm_smem / ctas - 1024; if (per_cta > dev_smem) per_cta = dev_smem;
          int k = (int)((per_cta - mb) / wb); if (k > 16) k = 16; if (k < 0) k = 0;
          kk[ctas] = k; if (k * ctas > best) best = k * ctas;
        }
        // the batch is worked off in rounds of (resident warps) environments: take the shape with the fewest rounds among those within 15 % of the best
        // residency (16 384 environments: one CTA of 16 warps does it in 7 rounds, two CTAs of 7 in 8; measured 28.6 vs 25.3 M env-steps/s); on a
        // tie the fewest CTAs per SM -- with the static schedule nothing is refilled mid-round, and one CTA of 14 warps (the balanced count below)
        // measured 1 % faster than two of 7 at 4096 environments (26.25 vs 25.96 M)
        long best_rounds = -1;
        for (int ctas = 1; ctas <= 4; ++ctas) if (kk[ctas] >= 1 && kk[ctas] * ctas * 100 >= best * 85) {
          const long slots = (long)kk[ctas] * ctas * (sms > 0 ? sms : 1), rounds = (n + slots - 1) / slots;
          if (best_rounds < 0 || rounds < best_rounds) { best_rounds = rounds; k_sel = kk[ctas]; }
        }
      }
      if (k_sel < 1 || k_sel > 16 || (long)model_bytes<real>() + k_sel * wb > (long)dev_smem) { set_err("not enough shared memory per block"); return false; }
      if (!w && sms > 0) {   // balance the rounds: with one resident CTA per SM, n / sms environments per CTA are worked off in ceil(. / k) rounds of
                             // k; the smallest k with the same number of rounds leaves fewer warps contending in every round
        int ctas_per_sm = 1; { long per_cta = (long)sm_smem / 2 - 1024; if ((per_cta - mb) / wb >= k_sel) ctas_per_sm = 2; }
        if (ctas_per_sm == 1) { const double per_cta = (double)n / sms; const int rounds = (int)std::ceil(per_cta / k_sel); if (rounds >= 1) { int kb = (int)std::ceil(per_cta / rounds); if (kb < 1) kb = 1; if (kb < k_sel) k_sel = kb; } }
      }
      cfg[ext].wpb = k_sel; cfg[ext].smem = model_bytes<real>() + (size_t)k_sel * wb;
      int per_sm = 0;
      { const void *entry = step_entry<real>(ext, feat);
        CUDA_OK(cudaFuncSetAttribute(entry, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg[ext].smem));
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, entry, 32 * k_sel, cfg[ext].smem)); }
      cfg[ext].resident_ctas = per_sm * sms < 1 ? 1 : per_sm * sms;
    }
    wpb = cfg[0].wpb;
    return reset(nullptr);
  }
  bool reset(const unsigned char *mask) override {
    CUDA_OK(cudaSetDevice(device));
    std::vector<real> qpos(QPOS_W_XB), qvel(QVEL_W_XB), qa(QVEL_W_XB), cst(CST_W), xf(XFRC_W); std::vector<int> df(DFILT_W);
    init_env_rows(hm, qpos.data(), qvel.data(), qa.data(), cst.data(), df.data(), xf.data());
    if (!mask) {
      std::vector<real> buf((size_t)n * CST_W); std::vector<int> ibuf((size_t)n * DFILT_W, 0);
      auto fill = [&](real *dst, const std::vector<real> &row, int w) -> bool { for (int e = 0; e < n; e++) memcpy(&buf[(size_t)e * w], row.data(), sizeof(real) * w); CUDA_OK(cudaMemcpyAsync(dst, buf.data(), sizeof(real) * n * w, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream)); return true; };
      if (!fill(A.qpos, qpos, QW) || !fill(A.qvel, qvel, VW) || !fill(A.qacc_ws, qa, VW) || !fill(A.cst, cst, CST_W) || !fill(A.xfrc, xf, XFRC_W)) return false;
      CUDA_OK(cudaMemcpyAsync(A.dfilt, ibuf.data(), sizeof(int) * n * DFILT_W, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream));
      return step(0, 1);
    }
    // masked reset: rewrite the selected rows, then mj_forward for exactly those environments
    if (!upload_mask(mask)) return false;
    if (!fill_masked(A.qpos, qpos.data(), QW) || !fill_masked(A.qvel, qvel.data(), VW) || !fill_masked(A.qacc_ws, qa.data(), VW) || !fill_masked(A.cst, cst.data(), CST_W) ||
        !fill_masked(A.xfrc, xf.data(), XFRC_W) || !fill_masked(A.dfilt, df.data(), DFILT_W)) return false;
    return step_masked(1);
  }
  bool upload_mask(const unsigned char *mask) {
    if (!d_mask) CUDA_OK(cudaMalloc(&d_mask, n));
    CUDA_OK(cudaMemcpyAsync(d_mask, mask, n, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    return true;
  }
  template <typename T> bool fill_masked(T *dst, const T *row, int w, int ncols = 1 << 30) {   // uses d_mask
    if (!d_row) CUDA_OK(cudaMalloc(&d_row, 4096));
    CUDA_OK(cudaMemcpyAsync(d_row, row, sizeof(T) * w, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    fill_rows_kernel<T><<<256, 256, 0, stream>>>(dst, (const T *)d_row, w, d_mask, n, ncols);
    CUDA_OK(cudaGetLastError()); CUDA_OK(cudaStreamSynchronize(stream));
    return true;
  }
  bool step_masked(int mode) { A.mask = d_mask; const bool ok = step(0, mode); A.mask = nullptr; return ok; }
  // ---- per-environment model constants (domain randomisation; reference setters src/cassiemujoco.c:1303-1436)
  bool ensure_cenv() {
    if (A.cenv) return true;
    CUDA_OK(cudaSetDevice(device));
    std::vector<real> row(CE_W), all((size_t)n * CE_W); init_cenv_row(h_model_copy, row.data(), hm, geom_dev);
    for (int e = 0; e < n; e++) memcpy(&all[(size_t)e * CE_W], row.data(), sizeof(real) * CE_W);
    CUDA_OK(cudaMalloc(&A.cenv, sizeof(real) * n * CE_W));
    CUDA_OK(cudaMemcpyAsync(A.cenv, all.data(), sizeof(real) * n * CE_W, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    return true;
  }
  int cenv_slot(const char *what, int i, int &width) const { return cassie::cenv_slot(hm, h_model_copy.nv, geom_dev, what, i, width); }
  bool set_model_rows(const char *what, const double *rows, int width) override {
    int w = 0; cenv_slot(what, 0, w);
    if (w < 0 || w != width) { set_err(std::string("set_model_rows: unknown array or wrong width for ") + what); return false; }
    if (!strcmp(what, "body_ipos") && h_model_copy.xb >= 0) for (int e = 0; e < n; e++) for (int k = 0; k < 3; k++)
      if (rows[(size_t)e * w + 3 * h_model_copy.xb + k] != hm.body_ipos[3 * h_model_copy.xb + k]) { set_err("body_ipos of the extra free body cannot be changed (its inertial frame must stay its body frame)"); return false; }
    if (!ensure_cenv()) return false;
    std::vector<real> all((size_t)n * CE_W);
    CUDA_OK(cudaMemcpyAsync(all.data(), A.cenv, sizeof(real) * n * CE_W, cudaMemcpyDeviceToHost, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    for (int e = 0; e < n; e++) for (int i = 0; i < w; i++) { int ww; const int sl = cenv_slot(what, i, ww); if (sl >= 0) all[(size_t)e * CE_W + sl] = (real)rows[(size_t)e * w + i]; }
    CUDA_OK(cudaMemcpyAsync(A.cenv, all.data(), sizeof(real) * n * CE_W, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    return true;
  }
  bool get_model_rows(const char *what, double *rows, int width) override {
    int w = 0; cenv_slot(what, 0, w);
    if (w < 0 || w != width) { set_err(std::string("get_model_rows: unknown array or wrong width for ") + what); return false; }
    if (!ensure_cenv()) return false;
    std::vector<real> all((size_t)n * CE_W);
    CUDA_OK(cudaMemcpyAsync(all.data(), A.cenv, sizeof(real) * n * CE_W, cudaMemcpyDeviceToHost, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    const std::vector<double> *src = !strcmp(what, "geom_friction") ? &hm.geom_friction : nullptr;
    for (int e = 0; e < n; e++) for (int i = 0; i < w; i++) { int ww; const int sl = cenv_slot(what, i, ww); rows[(size_t)e * w + i] = sl >= 0 ? (double)all[(size_t)e * CE_W + sl] : (src ? (*src)[i] : 0.0); }
    return true;
  }
  // mj_setConst for the selected environments on the device (mode 3), optionally followed by the state reset cassie_sim_set_const performs
  bool set_const(const unsigned char *mask, bool reset_state) override {
    if (!ensure_cenv()) return false;
    std::vector<unsigned char> ones; if (!mask) { ones.assign(n, 1); mask = ones.data(); }
    if (!upload_mask(mask) || !step_masked(3)) return false;
    if (!reset_state) return sync();
    // src/cassiemujoco.c:955-971: qpos <- the init constants, qvel <- 0, time <- 0, mj_forward (filters, delay line, cassie_out keep their values)
    std::vector<real> qpos(QPOS_W_XB), qvel(QVEL_W_XB), qa(QVEL_W_XB), cst(CST_W), xf(XFRC_W); std::vector<int> df(DFILT_W);
    init_env_rows(hm, qpos.data(), qvel.data(), qa.data(), cst.data(), df.data(), xf.data());
    if (!fill_masked(A.qpos, qpos.data(), QW, 35) || !fill_masked(A.qvel, qvel.data(), VW)) return false;   // the reference copies 35 qpos entries (:967): an extra free body keeps its pose
    { std::vector<real> all((size_t)n * CST_W);
      CUDA_OK(cudaMemcpyAsync(all.data(), A.cst, sizeof(real) * n * CST_W, cudaMemcpyDeviceToHost, stream)); CUDA_OK(cudaStreamSynchronize(stream));
      for (int e = 0; e < n; e++) if (mask[e]) all[(size_t)e * CST_W + CS_TIME] = 0;
      CUDA_OK(cudaMemcpyAsync(A.cst, all.data(), sizeof(real) * n * CST_W, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream)); }
    return step_masked(1) && sync();
  }
  bool set_pd(const double *pd) override {
    h_tmp.resize((size_t)n * PD_W);
    for (size_t i = 0; i < (size_t)n * PD_W; i++) h_tmp[i] = (real)pd[i];
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaMemcpyAsync(A.pd, h_tmp.data(), sizeof(real) * n * PD_W, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaStreamSynchronize(stream));   // h_tmp is pageable and reused
    return true;
  }
  // cassie_sim_step_pd for every env with host AoS buffers: pack pd_in_t[] -> pinned rows -> H2D, one tick, D2H rows -> state_out_t[]
  bool step_pd_aos(const pd_in_t *pd_in, state_out_t *state_out, const double *radio) override {
    CUDA_OK(cudaSetDevice(device));
    static const int aos_threads = aos_thread_count();
    static const bool aos_events = getenv("CASSIE_B200_AOS_EVENTS") != nullptr;
    static const bool aos_obs_dma = [] { const char *e = getenv("CASSIE_B200_AOS_OBS_DMA"); return e && atoi(e) != 0; }();
    static const int aos_chunks = [] { const char *e = getenv("CASSIE_B200_AOS_CHUNKS"); int t = e ? atoi(e) : 2; return t < 1 ? 1 : (t > AOS_MAXC ? AOS_MAXC : t); }();   // launches per call for batches of up to two rounds
    if (!pin_pd) {   // pinned AND mapped: the step kernel itself writes the observation rows into pin_obs
      CUDA_OK(cudaHostAlloc(&pin_pd, sizeof(real) * n * PD_W, cudaHostAllocMapped)); CUDA_OK(cudaHostAlloc(&pin_obs, sizeof(real) * n * OBS_W, cudaHostAllocMapped));
      CUDA_OK(cudaHostGetDevicePointer(&dpin_pd, pin_pd, 0)); CUDA_OK(cudaHostGetDevicePointer(&dpin_obs, pin_obs, 0));
      for (int i = 0; i < AOS_MAXC; i++) { CUDA_OK(cudaEventCreateWithFlags(&aos_done[i], cudaEventDisableTiming)); CUDA_OK(cudaStreamCreateWithFlags(&aos_stream[i], cudaStreamNonBlocking)); }
      CUDA_OK(cudaEventCreateWithFlags(&aos_start, cudaEventDisableTiming));
    }
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto sec = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    double t_pack = 0, t_wait = 0, t_unpack = 0;
    auto tq = now();
    // taskPd branch of pd_in_t: task rows are uploaded only while some environment uses that branch.  While task rows exist, the whole batch is
    // scanned before the first launch (and the rows are dropped again when no environment uses them); otherwise -- the common case, motor PD only --
    // the scan rides along with the chunks' pack loops, and the first chunk that finds a task entry installs the rows for itself and the later chunks
    // (the earlier chunks, by their own scan, had none).
    auto uses_task = [](const pd_in_t &u) { int any = 0; const pd_task_in_t *t[2] = {&u.leftLeg.taskPd, &u.rightLeg.taskPd};
      for (int sd = 0; sd < 2; sd++) for (int k = 0; k < 6; k++) any |= (t[sd]->torque[k] != 0 || t[sd]->pGain[k] != 0 || t[sd]->dGain[k] != 0); return any; };
    auto install_task_rows = [&](cudaStream_t on) -> bool {
      if (!pin_task) CUDA_OK(cudaMallocHost(&pin_task, sizeof(real) * n * TASK_W));
      if (!A.task) { CUDA_OK(cudaMalloc(&A.task, sizeof(real) * n * TASK_W)); task_from_aos = true; }
#pragma omp parallel for schedule(static) num_threads(aos_threads) if (n >= 512)
      for (int e = 0; e < n; e++) { real *row = pin_task + (size_t)e * TASK_W;
        for (int sd = 0; sd < 2; sd++) { const pd_task_in_t *t = sd ? &pd_in[e].rightLeg.taskPd : &pd_in[e].leftLeg.taskPd; real *r = row + 30 * sd;
          for (int k = 0; k < 6; k++) { r[k] = (real)t->torque[k]; r[6 + k] = (real)t->pTarget[k]; r[12 + k] = (real)t->dTarget[k]; r[18 + k] = (real)t->pGain[k]; r[24 + k] = (real)t->dGain[k]; } }
        for (int i = 60; i < TASK_W; i++) row[i] = 0; }
      CUDA_OK(cudaMemcpyAsync(A.task, pin_task, sizeof(real) * n * TASK_W, cudaMemcpyHostToDevice, on));
      return true; };
    bool scan_in_chunks = (A.task == nullptr);
    if (!scan_in_chunks) {
      int any = 0;
#pragma omp parallel for schedule(static) num_threads(aos_threads) reduction(| : any) if (n >= 512)
      for (int e = 0; e < n; e++) any |= uses_task(pd_in[e]);
      if (any) { if (!install_task_rows(stream)) return false; }
      else if (task_from_aos) { CUDA_OK(cudaStreamSynchronize(stream)); CUDA_OK(cudaFree(A.task)); A.task = nullptr; task_from_aos = false; }   // rows installed with cassie_batch_set_task_pd stay
    }
    // cassie_sim_step_pd runs state_output_step in every call (src/cassiemujoco.c:1156): so does the batched entry point, inside the kernel,
    // from the first call that asks for state_out_t rows (unless the caller switched it off or runs the host-side checker instead)
    if (state_out && est_auto && !A.est && !est_filter && !enable_estimator_device(true)) return false;
    if (est_filter && est_state.size() != (size_t)n) est_state.assign((size_t)n, cassie::EstimatorFilter());
    const bool dev_est = A.est != nullptr;   // the estimator runs inside the kernel: its outputs are columns of the observation row
    // The batch is stepped as up to AOS_MAXC independent launches ("chunks") on their own streams: the host packs chunk c + 1 while chunk c is copied in
    // and stepping, and unpacks chunk c while the later ones are stepping; the kernels of consecutive chunks overlap at their edges (a chunk's CTAs start
    // on the SMs the previous chunk's CTAs have left, while its last observation rows are still draining over PCIe).  Chunks are whole rounds of the
    // resident warps when the batch is worked off in more than two rounds, equal parts otherwise.
    const int inst = (A.cenv || A.aux || A.task || (A.dbg && !DBG_IN_ALL_INSTANCES)) ? 1 : (A.est ? 2 : 0), slots = cfg[inst].resident_ctas * cfg[inst].wpb;
    int nchunk = 1, c0[AOS_MAXC + 1] = {0};
    if (state_out && n >= 1024 && aos_chunks > 1) {
      if (n <= 2 * slots) { nchunk = aos_chunks; for (int c = 1; c < nchunk; c++) c0[c] = (int)((size_t)n * c / nchunk + 31) / 32 * 32; }
      else { const int rounds = (n + slots - 1) / slots; nchunk = rounds < AOS_MAXC ? rounds : AOS_MAXC; const int per = (rounds + nchunk - 1) / nchunk * slots;
        for (int c = 1; c < nchunk; c++) c0[c] = per * c < n ? per * c : n; }
    }
    c0[nchunk] = n;
    CUDA_OK(cudaEventRecord(aos_start, stream));   // everything enqueued on the batch's stream so far happens before the chunks
    t_pack += sec(tq, now());
    for (int c = 0; c < nchunk; c++) {
      const int e0 = c0[c], e1 = c0[c + 1], cnt = e1 - e0; if (cnt <= 0) { CUDA_OK(cudaEventRecord(aos_done[c], stream)); continue; }
      tq = now();
      int any_task = 0;
#pragma omp parallel for schedule(static) num_threads(aos_threads) reduction(| : any_task) if (n >= 512)
      for (int e = e0; e < e1; e++) {
        real *row = pin_pd + (size_t)e * PD_W; const pd_in_t *u = pd_in + e;
        for (int i = 0; i < 10; i++) {
          const pd_motor_in_t *p = i < 5 ? &u->leftLeg.motorPd : &u->rightLeg.motorPd; const int k = i % 5;
          row[i] = (real)p->torque[k]; row[10 + i] = (real)p->pTarget[k]; row[20 + i] = (real)p->dTarget[k]; row[30 + i] = (real)p->pGain[k]; row[40 + i] = (real)p->dGain[k];
        }
        row[50] = row[51] = 0;
        if (scan_in_chunks) any_task |= uses_task(*u);
      }
      const size_t off = (size_t)e0 * PD_W; cudaStream_t cs = nchunk > 1 ? aos_stream[c] : stream;
      if (nchunk > 1) CUDA_OK(cudaStreamWaitEvent(cs, aos_start, 0));
      if (any_task) {   // first task entry of this call: rows for the whole batch, in place before this chunk's (and every later chunk's) launch
        if (!install_task_rows(cs)) return false;
        CUDA_OK(cudaStreamSynchronize(cs)); scan_in_chunks = false;
      }
      t_pack += sec(tq, now());
      if (aos_events && c == 0) { if (!aos_ev[0]) for (int i = 0; i < 3; i++) CUDA_OK(cudaEventCreate(&aos_ev[i])); CUDA_OK(cudaEventRecord(aos_ev[0], cs)); }
      CUDA_OK(cudaMemcpyAsync(A.pd + off, pin_pd + off, sizeof(real) * cnt * PD_W, cudaMemcpyHostToDevice, cs));   // one DMA per launch: a burst of small PCIe reads from every warp at kernel start measured no faster
      if (aos_events && c == 0) CUDA_OK(cudaEventRecord(aos_ev[1], cs));
      A.pd_host = nullptr; A.obs_host = (state_out && !aos_obs_dma) ? dpin_obs : nullptr;   // observation rows return by zero-copy stores as each environment finishes ...
      const bool launched = step_range(1, 0, e0, cnt, cs);
      A.obs_host = nullptr;
      if (!launched) return false;
      if (state_out && aos_obs_dma) CUDA_OK(cudaMemcpyAsync(pin_obs + (size_t)e0 * OBS_W, A.obs + (size_t)e0 * OBS_W, sizeof(real) * cnt * OBS_W, cudaMemcpyDeviceToHost, cs));   // ... or by one DMA per chunk
      CUDA_OK(cudaEventRecord(aos_done[c], cs));
      if (nchunk > 1) CUDA_OK(cudaStreamWaitEvent(stream, aos_done[c], 0));   // later work on the batch's stream sees the stepped rows
      if (aos_events && c == nchunk - 1) CUDA_OK(cudaEventRecord(aos_ev[2], cs));
    }
    if (!state_out) return sync();
    for (int c = 0; c < nchunk; c++) {
      tq = now();
      CUDA_OK(cudaEventSynchronize(aos_done[c]));
      t_wait += sec(tq, now()); tq = now();
#pragma omp parallel for schedule(static) num_threads(aos_threads) if (n >= 512)
      for (int e = c0[c]; e < c0[c + 1]; e++) {
      // every field of state_out_t is written exactly once (no memset pass): what the reference's block never writes (externalMoment,
      // terrain.slope, battery.current) is zero, as in a freshly set-up state_output_t
      const real *o = pin_obs + (size_t)e * OBS_W; state_out_t *y = state_out + e; const real *eo = o + OB_EST_OUT;
      for (int i = 0; i < 3; i++) { y->pelvis.position[i] = dev_est ? (double)eo[EO_POS + i] : 0.0; y->pelvis.translationalVelocity[i] = dev_est ? (double)eo[EO_VEL + i] : 0.0;
                                    y->pelvis.externalForce[i] = dev_est ? (double)eo[EO_EXTF + i] : 0.0; y->pelvis.externalMoment[i] = 0;
                                    y->pelvis.rotationalVelocity[i] = o[OB_GYRO + i]; y->pelvis.translationalAcceleration[i] = o[OB_EST_ACC + i]; }
      for (int i = 0; i < 4; i++) y->pelvis.orientation[i] = o[OB_EST_QUAT + i];
      y->terrain.height = dev_est ? (double)eo[EO_TERRAIN] : 0.0; y->terrain.slope[0] = y->terrain.slope[1] = 0;
      for (int i = 0; i < 10; i++) { y->motor.position[i] = o[OB_MPOS + i]; y->motor.velocity[i] = o[OB_MVEL + i]; y->motor.torque[i] = o[OB_MTORQUE + i]; }
      for (int i = 0; i < 6; i++) { y->joint.position[i] = o[OB_JPOS + i]; y->joint.velocity[i] = o[OB_JVEL + i]; }
      for (int sd = 0; sd < 2; sd++) {   // decoded estimator: foot pose (pelvis frame), velocities (foot frame), toe / heel force
        state_foot_out_t *f = sd ? &y->rightFoot : &y->leftFoot; const real *fo = o + OB_FOOT + 13 * sd;
        for (int i = 0; i < 3; i++) { f->position[i] = fo[i]; f->footRotationalVelocity[i] = fo[7 + i]; f->footTranslationalVelocity[i] = fo[10 + i]; }
        for (int i = 0; i < 4; i++) f->orientation[i] = fo[3 + i];
        if (est_forces || est_filter) {   // host-side checker path (cassie_batch_enable_estimator_forces / _filter): the same functions the kernel runs
          const double ang[7] = {o[OB_MPOS + 5 * sd], o[OB_MPOS + 5 * sd + 1], o[OB_MPOS + 5 * sd + 2], o[OB_MPOS + 5 * sd + 3], o[OB_JPOS + 3 * sd], o[OB_JPOS + 3 * sd + 1], o[OB_MPOS + 5 * sd + 4]};
          const double qd[4] = {o[OB_QUAT], o[OB_QUAT + 1], o[OB_QUAT + 2], o[OB_QUAT + 3]};
          estimator_leg_force(sd, ang, qd, f->toeForce);
        } else for (int i = 0; i < 3; i++) f->toeForce[i] = dev_est ? (double)eo[EO_TOE + 3 * sd + i] : 0.0;
        for (int i = 0; i < 3; i++) f->heelForce[i] = f->toeForce[i];
      }
      if (est_filter) {   // host-side checker path: one 2 kHz filter call per step_pd call, overwrites the four filtered outputs
        cassie::EstimatorFilter &f = est_state[(size_t)e];
        f.step(y->pelvis.orientation, y->leftFoot.position, y->rightFoot.position, y->leftFoot.toeForce[2] + y->leftFoot.heelForce[2],
               y->rightFoot.toeForce[2] + y->rightFoot.heelForce[2], y->pelvis.translationalAcceleration,
               y->pelvis.position, y->pelvis.translationalVelocity, y->pelvis.externalForce, &y->terrain.height);
      }
      for (int i = 0; i < 16; i++) y->radio.channel[i] = radio[(size_t)e * 16 + i];
      memset(&y->radio.signalGood, 0, sizeof(double)); y->radio.signalGood = true;   // the bool and its padding
      y->battery.stateOfCharge = 1; y->battery.current = 0;
      }
      t_unpack += sec(tq, now());
    }
    aos_t[0] += t_pack; aos_t[1] += t_wait; aos_t[2] += t_unpack; aos_t[3] += 1;
    if (aos_events) { float a = 0, b = 0; cudaEventElapsedTime(&a, aos_ev[0], aos_ev[1]); cudaEventElapsedTime(&b, aos_ev[1], aos_ev[2]); aos_t[4] += a; aos_t[5] += b; }
    return true;
  }
  // K terrains of nrow*ncol normalised elevations; environment e stands on terrain e % K (cassie_sim_set_hfielddata, src/cassiemujoco.c:2076-2080)
  bool set_hfield(const float *data, int n_terrains) override {
    if (hm.nhfield != 1 || n_terrains < 1) { set_err("this model has no height field"); return false; }
    CUDA_OK(cudaSetDevice(device));
    const size_t cells = (size_t)hm.hfield_nrow[0] * hm.hfield_ncol[0];
    CUDA_OK(cudaStreamSynchronize(stream));
    if (d_hfield) CUDA_OK(cudaFree(d_hfield));
    CUDA_OK(cudaMalloc(&d_hfield, sizeof(float) * cells * n_terrains));
    if (data) CUDA_OK(cudaMemcpy(d_hfield, data, sizeof(float) * cells * n_terrains, cudaMemcpyHostToDevice));
    else CUDA_OK(cudaMemset(d_hfield, 0, sizeof(float) * cells * n_terrains));
    A.hfield = d_hfield; A.n_terrain = n_terrains; A.hfield_stride = cells;
    return true;
  }
  // derived-quantity rows [n][AUX_W]: allocated on first use; while present every launch fills them (a few percent of a step)
  bool enable_aux(bool on) override {
    CUDA_OK(cudaSetDevice(device));
    if (on && !A.aux) { CUDA_OK(cudaMalloc(&A.aux, sizeof(real) * n * AUX_W)); CUDA_OK(cudaMemsetAsync(A.aux, 0, sizeof(real) * n * AUX_W, stream)); }
    if (!on && A.aux) { CUDA_OK(cudaStreamSynchronize(stream)); CUDA_OK(cudaFree(A.aux)); A.aux = nullptr; }
    return true;
  }
  bool has_aux() const override { return A.aux != nullptr; }
  bool enable_estimator_device(bool on) override {
    CUDA_OK(cudaSetDevice(device));
    if (on && !A.est) {
      CUDA_OK(cudaMalloc(&A.est, sizeof(double) * n * EST_W));
      CUDA_OK(cudaMemsetAsync(A.est, 0, sizeof(double) * n * EST_W, stream));
    }
    if (!on && A.est) { CUDA_OK(cudaStreamSynchronize(stream)); CUDA_OK(cudaFree(A.est)); A.est = nullptr;
                        CUDA_OK(cudaMemset2DAsync(A.obs + OB_EST_OUT, sizeof(real) * OBS_W, 0, sizeof(real) * EO_W, n, stream)); }   // its columns of the observation row read zero again
    return true;
  }
  bool reset_estimator_device(const unsigned char *mask) override {   // started flag (and everything else) back to zero: the filters start again at the next tick
    if (!A.est) return true;
    CUDA_OK(cudaSetDevice(device));
    if (!mask) { CUDA_OK(cudaMemsetAsync(A.est, 0, sizeof(double) * n * EST_W, stream)); return true; }
    if (!upload_mask(mask)) return false;
    const std::vector<double> zero(EST_W, 0.0);
    return fill_masked(A.est, zero.data(), EST_W);
  }
  // task-space PD rows (pd_in_t taskPd of both legs); NULL switches the branch off again
  bool set_task_pd(const double *rows) override {
    CUDA_OK(cudaSetDevice(device));
    task_from_aos = false;   // the caller owns the rows from here on: the AoS entry point no longer drops them
    if (!rows) { if (A.task) { CUDA_OK(cudaStreamSynchronize(stream)); CUDA_OK(cudaFree(A.task)); A.task = nullptr; } return true; }
    if (!A.task) CUDA_OK(cudaMalloc(&A.task, sizeof(real) * n * TASK_W));
    h_tmp.assign((size_t)n * TASK_W, 0);
    for (int e = 0; e < n; e++) for (int i = 0; i < 60; i++) h_tmp[(size_t)e * TASK_W + i] = (real)rows[(size_t)e * 60 + i];
    CUDA_OK(cudaMemcpyAsync(A.task, h_tmp.data(), sizeof(real) * n * TASK_W, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    return true;
  }
  // ---- snapshots (the Snap object owns its device memory and knows its own sizes: it may outlive this batch)
  Snap *snap_alloc() override { Snap *k = snap_new(device, n, (int)sizeof(real), QW, VW); if (!k) set_err("snapshot: out of device memory"); return k; }
  bool snap_compatible(const Snap *k) { if (k && k->n == n && k->esz == (int)sizeof(real) && k->qw == QW && k->vw == VW) return true; set_err("snapshot: taken from a batch of another size, precision or model"); return false; }
  void *snap_src(int r) { void *t[Snap::NROWS] = {A.qpos, A.qvel, A.qacc_ws, A.cst, A.xfrc, A.obs, A.dfilt, A.est, A.aux}; return t[r]; }
  bool snap_get(Snap *k) override {
    if (!snap_compatible(k)) return false;
    CUDA_OK(cudaSetDevice(device));
    k->has_est = A.est != nullptr; k->has_aux = A.aux != nullptr;
    for (int r = 0; r < Snap::NROWS; r++) {
      if ((r == Snap::R_EST && !k->has_est) || (r == Snap::R_AUX && !k->has_aux)) continue;
      if (!k->rows[r]) CUDA_OK(cudaMalloc(&k->rows[r], k->bytes(r)));
      CUDA_OK(cudaMemcpyAsync(k->rows[r], snap_src(r), k->bytes(r), cudaMemcpyDeviceToDevice, stream));
    }
    return sync();
  }
  bool snap_set(const Snap *k, const unsigned char *mask) override {
    if (!snap_compatible(k)) return false;
    CUDA_OK(cudaSetDevice(device));
    const unsigned char *dm = nullptr;
    if (mask) { if (!upload_mask(mask)) return false; dm = d_mask; }
    if (k->has_est && !A.est && !enable_estimator_device(true)) return false;
    if (!k->has_est && A.est && !reset_estimator_device(mask)) return false;   // the snapshot was taken before the estimator ran: it starts afresh
    for (int r = 0; r < Snap::NROWS; r++) {
      if ((r == Snap::R_EST && !k->has_est) || (r == Snap::R_AUX && (!k->has_aux || !A.aux))) continue;
      if (!dm) { CUDA_OK(cudaMemcpyAsync(snap_src(r), k->rows[r], k->bytes(r), cudaMemcpyDeviceToDevice, stream)); continue; }
      const int w = k->width(r);
      if (k->elem(r) == 8) copy_rows_kernel<unsigned long long><<<256, 256, 0, stream>>>((unsigned long long *)snap_src(r), (const unsigned long long *)k->rows[r], w, dm, n);
      else copy_rows_kernel<unsigned int><<<256, 256, 0, stream>>>((unsigned int *)snap_src(r), (const unsigned int *)k->rows[r], w, dm, n);
      CUDA_OK(cudaGetLastError());
    }
    return sync();
  }
  bool rebuild_model() override {
    CUDA_OK(cudaSetDevice(device));
    DevModel<real> *hmodel = (DevModel<real> *)calloc(1, model_bytes<real>()); std::string err; BuildInfo info;
    if (!build_dev_model(hm, *hmodel, err, &info)) { free(hmodel); set_err(err); return false; }
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaMemcpy(d_model, hmodel, model_bytes<real>(), cudaMemcpyHostToDevice));
    h_model_copy = *hmodel; free(hmodel);
    return true;
  }
  bool copy_model_from(BatchBase *other) override {
    Batch<real> *src = dynamic_cast<Batch<real> *>(other);
    if (!src || src->n != n || src->hm.nq != hm.nq || src->hm.nbody != hm.nbody) { set_err("copy: the two simulators differ in model, size or precision"); return false; }
    CUDA_OK(cudaSetDevice(device));
    hm = src->hm;
    if (!rebuild_model()) return false;
    if (src->A.cenv) { if (!ensure_cenv()) return false; CUDA_OK(cudaMemcpyAsync(A.cenv, src->A.cenv, sizeof(real) * n * CE_W, cudaMemcpyDeviceToDevice, stream)); }
    else if (A.cenv) { CUDA_OK(cudaStreamSynchronize(stream)); CUDA_OK(cudaFree(A.cenv)); A.cenv = nullptr; }
    if (src->d_hfield) {
      const size_t cells = (size_t)hm.hfield_nrow[0] * hm.hfield_ncol[0] * src->A.n_terrain;
      std::vector<float> tmp(cells); CUDA_OK(cudaMemcpy(tmp.data(), src->d_hfield, sizeof(float) * cells, cudaMemcpyDeviceToHost));
      if (!set_hfield(tmp.data(), src->A.n_terrain)) return false;
    }
    return sync();
  }
  // open-loop sinusoidal gait on the motor-PD targets (BASELINE config 5 "random PD gaits"): evaluated in the kernel every control tick, so
  // multi-tick launches follow it without a host round trip; installing it restarts the tick clock
  bool set_pd_gait(const double *amp, const double *freq, const double *phase) override {
    CUDA_OK(cudaSetDevice(device));
    if (!amp || !freq || !phase) { if (A.gait) { CUDA_OK(cudaStreamSynchronize(stream)); CUDA_OK(cudaFree(A.gait)); A.gait = nullptr; } return true; }
    if (!A.gait) CUDA_OK(cudaMalloc(&A.gait, sizeof(real) * n * GAIT_W));
    h_tmp.assign((size_t)n * GAIT_W, 0);
    for (int e = 0; e < n; e++) { real *r = &h_tmp[(size_t)e * GAIT_W]; for (int i = 0; i < 10; i++) { r[GA_AMP + i] = (real)amp[(size_t)e * 10 + i]; r[GA_PHASE + i] = (real)phase[(size_t)e * 10 + i]; } r[GA_FREQ] = (real)freq[e]; }
    CUDA_OK(cudaMemcpyAsync(A.gait, h_tmp.data(), sizeof(real) * n * GAIT_W, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemset2DAsync(A.dfilt + DF_TICK, sizeof(int) * DFILT_W, 0, sizeof(int), n, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    return true;
  }
  bool step(int nticks, int mode) override { return step_range(nticks, mode, 0, n, stream); }
  bool step_range(int nticks, int mode, int env0, int count, cudaStream_t on) {   // the launch covers environments [env0, env0 + count)
    CUDA_OK(cudaSetDevice(device));
    if (mode == 2 && !A.aux) { set_err("query needs the derived-quantity rows (cassie_batch_enable_aux)"); return false; }
    A.nsub = nsub_override;
    // per-environment model constants / derived-quantity rows / task-space PD / set_const: extended instance (its own scratch size and launch shape);
    // the in-kernel estimator alone: the plain instance with the estimator stage
    const int ext = (A.cenv || A.aux || A.task || (A.dbg && !DBG_IN_ALL_INSTANCES) || mode == 3) ? 1 : (A.est ? 2 : 0);   // debug batches: the stage dumps live in the extended instance
    A.warp_stride = (int)warp_bytes<real>(A.ystride, ext == 1);
    const LaunchCfg &c = cfg[ext];
    int grid = (count + c.wpb - 1) / c.wpb; if (grid > c.resident_ctas) grid = c.resident_ctas;
    EnvArrays<real> La = A; La.env0 = env0; La.n = count;   // the argument block of this launch
    { const DevModel<real> *dm = d_model; void *args[4] = {(void *)&dm, (void *)&La, (void *)&nticks, (void *)&mode};
      CUDA_OK(cudaLaunchKernel(step_entry<real>(ext, feat), dim3(grid), dim3(32 * c.wpb), args, c.smem, on)); }
    launches++;
    CUDA_OK(cudaGetLastError());
    return true;
  }
  bool integrate() override {
    CUDA_OK(cudaSetDevice(device));
    const int ntiles = (n + ITILE - 1) / ITILE; int sms = 0; CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const size_t ib = ISTAGES * itile_bytes<real>(A.qpos_w, A.qvel_w);
    CUDA_OK(cudaFuncSetAttribute(cassie_integrate_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ib));
    const int per_sm = (int)(220000 / ib) < 1 ? 1 : (int)(220000 / ib);
    const int grid = ntiles < sms * per_sm ? ntiles : sms * per_sm;
    cassie_integrate_kernel<real><<<grid, 256, ib, stream>>>(d_model, A.qpos, A.qvel, n, A.qpos_w, A.qvel_w);
    launches++;
    CUDA_OK(cudaGetLastError());
    return true;
  }
  bool d2h(const real *src, int w, int take, int off, double *out) {
    h_tmp.resize((size_t)n * w);
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaMemcpyAsync(h_tmp.data(), src, sizeof(real) * n * w, cudaMemcpyDeviceToHost, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    for (int e = 0; e < n; e++) for (int i = 0; i < take; i++) out[(size_t)e * take + i] = (double)h_tmp[(size_t)e * w + off + i];
    return true;
  }
  bool get(const char *f, double *out) override {
    if (!strcmp(f, "qpos")) return d2h(A.qpos, QW, hm.nq, 0, out);
    if (!strcmp(f, "qvel")) return d2h(A.qvel, VW, hm.nv, 0, out);
    if (!strcmp(f, "qacc_ws")) return d2h(A.qacc_ws, VW, hm.nv, 0, out);
    if (!strcmp(f, "time")) return d2h(A.cst, CST_W, 1, CS_TIME, out);
    if (!strcmp(f, "obs")) return d2h(A.obs, OBS_W, OBS_W, 0, out);
    if (!strcmp(f, "cst")) return d2h(A.cst, CST_W, CST_W, 0, out);
    if (!strcmp(f, "xfrc")) return d2h(A.xfrc, XFRC_W, XFRC_W, 0, out);
    if (!strcmp(f, "est_out")) { if (!A.est) { set_err("the in-kernel estimator is not enabled (cassie_batch_enable_estimator_device)"); return false; } return d2h(A.obs, OBS_W, EO_W, OB_EST_OUT, out); }
    if (!strcmp(f, "aux")) { if (!A.aux) { set_err("derived quantities are not enabled (cassie_batch_enable_aux)"); return false; } return d2h(A.aux, AUX_W, AUX_W, 0, out); }
    set_err(std::string("unknown field ") + f); return false;
  }
  bool h2d(real *dst, int w, int take, int off, const double *in) {
    h_tmp.resize((size_t)n * w);
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaMemcpyAsync(h_tmp.data(), dst, sizeof(real) * n * w, cudaMemcpyDeviceToHost, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    for (int e = 0; e < n; e++) for (int i = 0; i < take; i++) h_tmp[(size_t)e * w + off + i] = (real)in[(size_t)e * take + i];
    CUDA_OK(cudaMemcpyAsync(dst, h_tmp.data(), sizeof(real) * n * w, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    return true;
  }
  bool set(const char *f, const double *in) override {
    if (!strcmp(f, "qpos")) return h2d(A.qpos, QW, hm.nq, 0, in);
    if (!strcmp(f, "qvel")) return h2d(A.qvel, VW, hm.nv, 0, in);
    if (!strcmp(f, "time")) return h2d(A.cst, CST_W, 1, CS_TIME, in);
    if (!strcmp(f, "sto")) return h2d(A.cst, CST_W, 1, CS_STO, in);
    if (!strcmp(f, "cst")) return h2d(A.cst, CST_W, CST_W, 0, in);
    set_err(std::string("unknown field ") + f); return false;
  }
  bool set_xfrc(const double *xfrc, int body) override {
    h_tmp.assign((size_t)n * XFRC_W, 0);
    if (xfrc) for (int e = 0; e < n; e++) { for (int i = 0; i < 6; i++) h_tmp[(size_t)e * XFRC_W + i] = (real)xfrc[(size_t)e * 6 + i]; h_tmp[(size_t)e * XFRC_W + 6] = (real)body; }
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaMemcpyAsync(A.xfrc, h_tmp.data(), sizeof(real) * n * XFRC_W, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    return true;
  }
  void *dev_ptr(const char *f) override {
    if (!strcmp(f, "qpos")) return A.qpos; if (!strcmp(f, "qvel")) return A.qvel; if (!strcmp(f, "pd")) return A.pd; if (!strcmp(f, "obs")) return A.obs;
    if (!strcmp(f, "xfrc")) return A.xfrc; if (!strcmp(f, "cst")) return A.cst; if (!strcmp(f, "qacc_ws")) return A.qacc_ws; if (!strcmp(f, "aux")) return A.aux; if (!strcmp(f, "task")) return A.task;
    return nullptr;
  }
  bool get_counters(int *out) override {
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaMemcpyAsync(out, A.counters, sizeof(int) * n * 8, cudaMemcpyDeviceToHost, stream)); CUDA_OK(cudaStreamSynchronize(stream));
    return true;
  }
  int debug_dump(int env, double *out, int cnt) override {
    if (!A.dbg || env < 0 || env >= n) return -1;
    std::vector<real> t(D_SIZE);
    cudaSetDevice(device);
    if (cudaMemcpyAsync(t.data(), A.dbg + (size_t)env * D_SIZE, sizeof(real) * D_SIZE, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) return -1;
    if (cnt > D_SIZE) cnt = D_SIZE;
    for (int i = 0; i < cnt; i++) out[i] = (double)t[i];
    return cnt;
  }
};

}  // namespace cassie

// ====================================================================== C-ABI
using namespace cassie;
struct cassie_batch { BatchBase *impl; std::vector<double> obs, radio; };
struct cassie_sim { cassie_batch *b; double qpos[64], qvel[64], time, qpos_dev[64], qvel_dev[64], time_dev, aux[AUX_W]; std::vector<float> hfield, hfield_dev; std::string path; double timestep = 0, timestep_dev = 0;
  std::vector<double> m_mass, m_ipos, m_damp, m_fric, m_mass_dev, m_ipos_dev, m_damp_dev, m_fric_dev;
  std::vector<double> m_gpos, m_gquat, m_gsize, m_gpos_dev, m_gquat_dev, m_gsize_dev; };   // geom placement (stair boxes): host mirrors, the constant block is rebuilt when they change   // host mirrors of the model arrays handed out as borrowed pointers

// cassie_state_t (include/cassiemujoco.h:434-463): a device-resident snapshot of one simulator's rows plus the host mirrors the reference hands out as
// borrowed pointers (time, qpos, qvel); the snapshot is allocated by the first cassie_get_state, which also fixes the owning simulator's sizes
struct cassie_state { cassie::Snap *snap = nullptr; double qpos[64], qvel[64], time = 0, qpos_snap[64], qvel_snap[64], time_snap = 0; };

static std::mutex g_model_mutex;
static std::string g_model_path;   // what cassie_mujoco_init cached (the reference caches the parsed model, src/cassiemujoco.c:48-59)

extern "C" {

const char *cassie_b200_last_error(void) { return g_err.c_str(); }

cassie_batch_t *cassie_batch_init(const char *modelfile, int n_env, int device, int precision) {
  g_err.clear();
  if (!modelfile || n_env <= 0) { set_err("cassie_batch_init: bad arguments"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_err("no CUDA device: the batched stepper has no CPU fallback"); return nullptr; }
  if (device < 0 || device >= ndev) { set_err("bad device index"); return nullptr; }
  BatchBase *impl = precision == CASSIE_B200_FP64 ? (BatchBase *)new Batch<double>() : (BatchBase *)new Batch<float>();
  std::string err;
  if (!load_model_any(modelfile, impl->hm, err)) { set_err("Load model error: " + err); delete impl; return nullptr; }
  impl->n = n_env; impl->device = device; impl->precision = precision; impl->debug = getenv("CASSIE_B200_DEBUG") != nullptr;
  if (!impl->init() || !impl->sync()) { delete impl; return nullptr; }
  cassie_batch *b = new cassie_batch(); b->impl = impl; b->radio.assign((size_t)n_env * 16, 0.0);
  for (int e = 0; e < n_env; e++) b->radio[(size_t)e * 16 + 8] = 1.0;
  return b;
}
void cassie_batch_free(cassie_batch_t *b) { if (!b) return; delete b->impl; delete b; }
int cassie_batch_nenv(const cassie_batch_t *b) { return b->impl->n; }
int cassie_batch_nq(const cassie_batch_t *b) { return b->impl->hm.nq; }
int cassie_batch_nv(const cassie_batch_t *b) { return b->impl->hm.nv; }
int cassie_batch_precision(const cassie_batch_t *b) { return b->impl->precision; }
int cassie_batch_row_width(const cassie_batch_t *b, const char *field) {
  if (!strcmp(field, "qpos")) return b->impl->hm.nq > 36 ? QPOS_W_XB : QPOS_W_MAIN; if (!strcmp(field, "qvel")) return b->impl->hm.nv > 32 ? QVEL_W_XB : QVEL_W_MAIN;
  if (!strcmp(field, "pd")) return PD_W; if (!strcmp(field, "obs")) return OBS_W; if (!strcmp(field, "xfrc")) return XFRC_W; if (!strcmp(field, "aux")) return AUX_W; return -1;
}
long cassie_batch_launch_count(const cassie_batch_t *b) { return b->impl->launches; }
void cassie_batch_reset(cassie_batch_t *b, const unsigned char *mask) { b->impl->reset(mask); b->impl->sync(); b->impl->reset_estimator(mask); b->impl->reset_estimator_device(mask); }   // a fresh cassie_sim_t has a fresh estimator
void cassie_batch_set_pd(cassie_batch_t *b, const double *pd) { b->impl->set_pd(pd); }
void cassie_batch_step(cassie_batch_t *b, int nticks) { if (nticks > 0) b->impl->step(nticks, 0); }
void cassie_batch_forward(cassie_batch_t *b) { b->impl->step(0, 1); }
void cassie_batch_sync(cassie_batch_t *b) { b->impl->sync(); }
void cassie_batch_get_qpos(cassie_batch_t *b, double *out) { b->impl->get("qpos", out); }
void cassie_batch_set_qpos(cassie_batch_t *b, const double *in) { b->impl->set("qpos", in); }
void cassie_batch_get_qvel(cassie_batch_t *b, double *out) { b->impl->get("qvel", out); }
void cassie_batch_set_qvel(cassie_batch_t *b, const double *in) { b->impl->set("qvel", in); }
void cassie_batch_get_time(cassie_batch_t *b, double *out) { b->impl->get("time", out); }
void cassie_batch_get_obs(cassie_batch_t *b, double *out) { b->impl->get("obs", out); }
int cassie_batch_nbody(const cassie_batch_t *b) { return b->impl->hm.nbody; }
int cassie_batch_ngeom(const cassie_batch_t *b) { return b->impl->hm.ngeom; }
int cassie_batch_set_body_mass(cassie_batch_t *b, const double *mass) { return b->impl->set_model_rows("body_mass", mass, b->impl->hm.nbody) ? 0 : -1; }
int cassie_batch_set_body_ipos(cassie_batch_t *b, const double *ipos) { return b->impl->set_model_rows("body_ipos", ipos, 3 * b->impl->hm.nbody) ? 0 : -1; }
int cassie_batch_set_dof_damping(cassie_batch_t *b, const double *damp) { return b->impl->set_model_rows("dof_damping", damp, b->impl->hm.nv) ? 0 : -1; }
int cassie_batch_set_geom_friction(cassie_batch_t *b, const double *fric) { return b->impl->set_model_rows("geom_friction", fric, 3 * b->impl->hm.ngeom) ? 0 : -1; }
int cassie_batch_get_body_mass(cassie_batch_t *b, double *mass) { return b->impl->get_model_rows("body_mass", mass, b->impl->hm.nbody) ? 0 : -1; }
int cassie_batch_get_body_ipos(cassie_batch_t *b, double *ipos) { return b->impl->get_model_rows("body_ipos", ipos, 3 * b->impl->hm.nbody) ? 0 : -1; }
int cassie_batch_get_dof_damping(cassie_batch_t *b, double *damp) { return b->impl->get_model_rows("dof_damping", damp, b->impl->hm.nv) ? 0 : -1; }
int cassie_batch_get_geom_friction(cassie_batch_t *b, double *fric) { return b->impl->get_model_rows("geom_friction", fric, 3 * b->impl->hm.ngeom) ? 0 : -1; }
int cassie_batch_set_const(cassie_batch_t *b, const unsigned char *mask, int reset_state) { return b->impl->set_const(mask, reset_state != 0) ? 0 : -1; }
int cassie_batch_set_task_pd(cassie_batch_t *b, const double *rows) { return b->impl->set_task_pd(rows) ? 0 : -1; }
int cassie_batch_set_pd_gait(cassie_batch_t *b, const double *amp, const double *freq, const double *phase) { return b->impl->set_pd_gait(amp, freq, phase) ? 0 : -1; }
int cassie_batch_enable_estimator_forces(cassie_batch_t *b, int on) { b->impl->est_forces = on != 0; return 0; }
void cassie_b200_estimator_leg_force(int side, const double ang[7], const double quat[4], double force[3]) { estimator_leg_force(side, ang, quat, force); }
int cassie_batch_enable_estimator_filter(cassie_batch_t *b, int on) { b->impl->est_filter = on != 0; if (on) b->impl->est_forces = true; else b->impl->est_state.clear(); return 0; }
int cassie_batch_reset_estimator(cassie_batch_t *b, const unsigned char *mask) { b->impl->reset_estimator(mask); return b->impl->reset_estimator_device(mask) ? 0 : -1; }
// the same filters as a plain host object (no GPU involved): what the unpack loop above runs per environment
void *cassie_b200_estimator_filter_new(void) { return new cassie::EstimatorFilter(); }
void cassie_b200_estimator_filter_free(void *f) { delete static_cast<cassie::EstimatorFilter *>(f); }
void cassie_b200_estimator_filter_reset(void *f) { static_cast<cassie::EstimatorFilter *>(f)->reset(); }
void cassie_b200_estimator_filter_step(void *f, state_out_t *y) {
  static_cast<cassie::EstimatorFilter *>(f)->step(y->pelvis.orientation, y->leftFoot.position, y->rightFoot.position, y->leftFoot.toeForce[2] + y->leftFoot.heelForce[2],
      y->rightFoot.toeForce[2] + y->rightFoot.heelForce[2], y->pelvis.translationalAcceleration, y->pelvis.position, y->pelvis.translationalVelocity, y->pelvis.externalForce, &y->terrain.height);
}
int cassie_batch_enable_aux(cassie_batch_t *b, int on) { return b->impl->enable_aux(on != 0) ? 0 : -1; }
int cassie_batch_get_aux(cassie_batch_t *b, double *out) { return b->impl->get("aux", out) ? 0 : -1; }
int cassie_batch_enable_estimator_device(cassie_batch_t *b, int on) { if (!on) b->impl->est_auto = false; return b->impl->enable_estimator_device(on != 0) ? 0 : -1; }
int cassie_batch_get_estimator(cassie_batch_t *b, double *out) { return b->impl->get("est_out", out) ? 0 : -1; }
int cassie_batch_query(cassie_batch_t *b) { if (!b->impl->has_aux() && !b->impl->enable_aux(true)) return -1; return b->impl->step(0, 2) ? 0 : -1; }
int cassie_batch_apply_force(cassie_batch_t *b, const double *xfrc, const char *body_name) {
  int id = body_name ? b->impl->hm.body_id(body_name) : -1;
  if (id <= 0) { set_err("cassie_batch_apply_force: unknown body (no-op)"); return -1; }
  return b->impl->set_xfrc(xfrc, id) ? 0 : -1;
}
void cassie_batch_clear_forces(cassie_batch_t *b) { b->impl->set_xfrc(nullptr, 0); }
void cassie_batch_integrate_pos(cassie_batch_t *b) { b->impl->integrate(); }
int cassie_batch_set_hfielddata(cassie_batch_t *b, const float *data, int n_terrains) { return b->impl->set_hfield(data, n_terrains) ? 0 : -1; }
int cassie_batch_hfield_nrow(const cassie_batch_t *b) { return b->impl->hm.nhfield ? b->impl->hm.hfield_nrow[0] : 0; }
int cassie_batch_hfield_ncol(const cassie_batch_t *b) { return b->impl->hm.nhfield ? b->impl->hm.hfield_ncol[0] : 0; }
void *cassie_batch_device_ptr(cassie_batch_t *b, const char *field) { return b->impl->dev_ptr(field); }
void cassie_batch_set_stream(cassie_batch_t *b, void *s) { b->impl->sync(); if (b->impl->own_stream && b->impl->stream) cudaStreamDestroy(b->impl->stream); b->impl->stream = (cudaStream_t)s; b->impl->own_stream = false; }
void *cassie_batch_get_stream(cassie_batch_t *b) { return (void *)b->impl->stream; }
void cassie_batch_get_counters(cassie_batch_t *b, int *out) { b->impl->get_counters(out); }
int cassie_batch_debug_dump(cassie_batch_t *b, int env, double *out, int n) { return b->impl->debug_dump(env, out, n); }

void cassie_batch_aos_timing(cassie_batch_t *b, double out[6], int reset) { for (int i = 0; i < 6; i++) out[i] = b->impl->aos_t[i]; if (reset) for (int i = 0; i < 6; i++) b->impl->aos_t[i] = 0; }
int cassie_b200_aos_threads(void) { return cassie::aos_thread_count(); }
int cassie_b200_effective_cpus(void) { return cassie::effective_cpus(); }
void cassie_sim_step_pd_batch(cassie_batch_t *b, const pd_in_t *pd_in, state_out_t *state_out) {
  b->impl->step_pd_aos(pd_in, state_out, b->radio.data());
}

// ---------------- legacy single-environment verbs
bool cassie_mujoco_init(const char *modelfile) {
  std::lock_guard<std::mutex> g(g_model_mutex);
  if (!g_model_path.empty()) return true;
  if (!modelfile) { set_err("cassie_mujoco_init: model file required"); return false; }
  HostModel hm; std::string err;
  if (!load_model_any(modelfile, hm, err)) { fprintf(stderr, "Load model error: %s\n", err.c_str()); g_err = err; return false; }
  g_model_path = modelfile; return true;
}
void cassie_cleanup(void) { std::lock_guard<std::mutex> g(g_model_mutex); g_model_path.clear(); }
static void sim_pull(cassie_sim_t *c) {
  c->b->impl->get("qpos", c->qpos); c->b->impl->get("qvel", c->qvel); c->b->impl->get("time", &c->time); if (c->b->impl->has_aux()) c->b->impl->get("aux", c->aux);
  memcpy(c->qpos_dev, c->qpos, sizeof c->qpos); memcpy(c->qvel_dev, c->qvel, sizeof c->qvel); c->time_dev = c->time;
}
static void sim_push(cassie_sim_t *c) {  // upload whatever the caller wrote through the borrowed pointers
  if (memcmp(c->qpos_dev, c->qpos, sizeof c->qpos)) c->b->impl->set("qpos", c->qpos);
  if (memcmp(c->qvel_dev, c->qvel, sizeof c->qvel)) c->b->impl->set("qvel", c->qvel);
  if (c->time_dev != c->time) c->b->impl->set("time", &c->time);
  if (c->timestep != c->timestep_dev) { if (c->timestep > 0) { c->b->impl->hm.timestep = c->timestep; c->b->impl->rebuild_model(); } c->timestep_dev = c->timestep; }   // written through cassie_sim_timestep()
  if (!c->hfield.empty() && c->hfield != c->hfield_dev) { c->b->impl->set_hfield(c->hfield.data(), 1); c->hfield_dev = c->hfield; }
  if (c->m_mass != c->m_mass_dev) { cassie_batch_set_body_mass(c->b, c->m_mass.data()); c->m_mass_dev = c->m_mass; }
  if (c->m_ipos != c->m_ipos_dev) { cassie_batch_set_body_ipos(c->b, c->m_ipos.data()); c->m_ipos_dev = c->m_ipos; }
  if (c->m_damp != c->m_damp_dev) { cassie_batch_set_dof_damping(c->b, c->m_damp.data()); c->m_damp_dev = c->m_damp; }
  if (c->m_fric != c->m_fric_dev) { cassie_batch_set_geom_friction(c->b, c->m_fric.data()); c->m_fric_dev = c->m_fric; }
  if (c->m_gpos != c->m_gpos_dev || c->m_gquat != c->m_gquat_dev || c->m_gsize != c->m_gsize_dev) {   // a geom was moved / turned / resized (src/cassiemujoco.c:1438-1541)
    HostModel &hm = c->b->impl->hm; hm.geom_pos = c->m_gpos; hm.geom_quat = c->m_gquat; hm.geom_size = c->m_gsize;
    c->b->impl->rebuild_model(); c->m_gpos_dev = c->m_gpos; c->m_gquat_dev = c->m_gquat; c->m_gsize_dev = c->m_gsize; }
}
cassie_sim_t *cassie_sim_init(const char *modelfile, bool reinit) {
  std::string path;
  { std::lock_guard<std::mutex> g(g_model_mutex); if (reinit || g_model_path.empty()) { if (!modelfile) { set_err("cassie_sim_init: model file required"); return nullptr; } path = modelfile; if (g_model_path.empty()) g_model_path = modelfile; } else path = g_model_path; }
  cassie_batch_t *b = cassie_batch_init(path.c_str(), 1, 0, CASSIE_B200_FP64);
  if (!b) return nullptr;
  cassie_sim_t *c = new cassie_sim(); memset(c->qpos, 0, sizeof c->qpos); memset(c->qvel, 0, sizeof c->qvel); memset(c->aux, 0, sizeof c->aux); c->b = b; c->path = path; c->timestep = c->timestep_dev = b->impl->hm.timestep;
  b->impl->enable_aux(true); b->impl->step(0, 1);   // a single environment always carries the derived-quantity row; populate it like the sensordata
  b->impl->enable_estimator_device(true);   // ... and the estimator (toe / heel forces, filters) inside the kernel, as state_output_step runs in every cassie_sim_step_pd (src/cassiemujoco.c:1156)
  sim_pull(c);
  { const HostModel &hm = b->impl->hm; c->m_mass = hm.body_mass; c->m_ipos = hm.body_ipos; c->m_damp = hm.dof_damping; c->m_fric = hm.geom_friction;
    c->m_mass_dev = c->m_mass; c->m_ipos_dev = c->m_ipos; c->m_damp_dev = c->m_damp; c->m_fric_dev = c->m_fric;
    c->m_gpos = hm.geom_pos; c->m_gquat = hm.geom_quat; c->m_gsize = hm.geom_size; c->m_gpos_dev = c->m_gpos; c->m_gquat_dev = c->m_gquat; c->m_gsize_dev = c->m_gsize; }
  if (b->impl->hm.nhfield) { c->hfield.assign((size_t)b->impl->hm.hfield_nrow[0] * b->impl->hm.hfield_ncol[0], 0.0f); c->hfield_dev = c->hfield; }
  return c;
}
void cassie_sim_free(cassie_sim_t *c) { if (!c) return; cassie_batch_free(c->b); delete c; }
void cassie_sim_step_pd(cassie_sim_t *c, state_out_t *y, const pd_in_t *u) { sim_push(c); cassie_sim_step_pd_batch(c->b, u, y); sim_pull(c); }
double *cassie_sim_time(cassie_sim_t *c) { return &c->time; }
double *cassie_sim_qpos(cassie_sim_t *c) { return c->qpos; }
double *cassie_sim_qvel(cassie_sim_t *c) { return c->qvel; }
int cassie_sim_nv(const cassie_sim_t *c) { return c->b->impl->hm.nv; }
int cassie_sim_nq(const cassie_sim_t *c) { return c->b->impl->hm.nq; }
void cassie_sim_apply_force(cassie_sim_t *c, double xfrc[6], const char *name) { cassie_batch_apply_force(c->b, xfrc, name); }
void cassie_sim_clear_forces(cassie_sim_t *c) { cassie_batch_clear_forces(c->b); }
void cassie_sim_radio(cassie_sim_t *c, double channels[16]) { for (int i = 0; i < 16; i++) c->b->radio[i] = channels[i]; double s = channels[8]; c->b->impl->set("sto", &s); }
int cassie_sim_get_hfield_nrow(cassie_sim_t *c) { return cassie_batch_hfield_nrow(c->b); }
int cassie_sim_get_hfield_ncol(cassie_sim_t *c) { return cassie_batch_hfield_ncol(c->b); }
int cassie_sim_get_nhfielddata(cassie_sim_t *c) { return (int)c->hfield.size(); }
float *cassie_sim_hfielddata(cassie_sim_t *c) { return c->hfield.empty() ? nullptr : c->hfield.data(); }
void cassie_sim_set_hfielddata(cassie_sim_t *c, float *data) { for (size_t i = 0; i < c->hfield.size(); i++) c->hfield[i] = data[i]; }
// ---- read-only derived quantities (include/cassiemujoco.h:200-240).  Contact forces, flags, foot positions / velocities are by-products of
// the last step (the reference reads the same stale mjData arrays); the centre-of-mass group recomputes the kinematics of the CURRENT
// state first, as the reference does with mj_fwdPosition, through a query launch that writes nothing else.
static const double *sim_aux(const cassie_sim_t *c) { return c->aux; }
static void sim_query(const cassie_sim_t *cc) { cassie_sim_t *c = const_cast<cassie_sim_t *>(cc); sim_push(c); double keep[AUX_W]; memcpy(keep, c->aux, sizeof keep);
  cassie_batch_query(c->b); c->b->impl->get("aux", c->aux);
  for (int i = 0; i < AUX_W; i++) if (i < AX_CM_POS || i >= AX_ANGMOM + 3) c->aux[i] = keep[i]; }
void cassie_sim_foot_forces(const cassie_sim_t *c, double cfrc[12]) { memcpy(cfrc, sim_aux(c) + AX_FOOT_FORCE, 12 * sizeof(double)); }
void cassie_sim_heeltoe_forces(const cassie_sim_t *c, double toe_force[6], double heel_force[6]) { memcpy(toe_force, sim_aux(c) + AX_TOE_FORCE, 6 * sizeof(double)); memcpy(heel_force, sim_aux(c) + AX_HEEL_FORCE, 6 * sizeof(double)); }
void cassie_sim_foot_positions(const cassie_sim_t *c, double cpos[6]) { memcpy(cpos, sim_aux(c) + AX_FOOT_POS, 6 * sizeof(double)); }
void cassie_sim_foot_velocities(const cassie_sim_t *c, double cvel[12]) { memcpy(cvel, sim_aux(c) + AX_FOOT_VEL, 12 * sizeof(double)); }
void cassie_sim_cm_position(const cassie_sim_t *c, double cm_pos[3]) { sim_query(c); memcpy(cm_pos, sim_aux(c) + AX_CM_POS, 3 * sizeof(double)); }
void cassie_sim_cm_velocity(const cassie_sim_t *c, double cm_vel[3]) { sim_query(c); memcpy(cm_vel, sim_aux(c) + AX_CM_VEL, 3 * sizeof(double)); }
void cassie_sim_angular_momentum(const cassie_sim_t *c, double Lcm[3]) { sim_query(c); memcpy(Lcm, sim_aux(c) + AX_ANGMOM, 3 * sizeof(double)); }
bool cassie_sim_check_obstacle_collision(const cassie_sim_t *c) { return sim_aux(c)[AX_OBSTACLE] != 0; }
bool cassie_sim_check_self_collision(const cassie_sim_t *c) { return sim_aux(c)[AX_SELF] != 0; }
bool cassie_sim_geom_collision(const cassie_sim_t *c, int geom_group) { return geom_group >= 0 && geom_group < 16 && (((int)sim_aux(c)[AX_GROUPMASK] >> geom_group) & 1); }
// ---- model constants (domain randomisation).  include/cassiemujoco.h / src/cassiemujoco.c:1303-1436: the reference hands out pointers into
// its private mjModel copy; here they point into host mirrors that are uploaded before the next step / query / set_const.
void cassie_sim_params(cassie_sim_t *c, int *params) { const HostModel &hm = c->b->impl->hm; params[0] = hm.nq; params[1] = hm.nv; params[2] = hm.nu; params[3] = 29; params[4] = hm.nbody; params[5] = hm.ngeom; }
double *cassie_sim_dof_damping(cassie_sim_t *c) { return c->m_damp.data(); }
double *cassie_sim_body_mass(cassie_sim_t *c) { return c->m_mass.data(); }
double *cassie_sim_body_ipos(cassie_sim_t *c) { return c->m_ipos.data(); }
double *cassie_sim_geom_friction(cassie_sim_t *c) { return c->m_fric.data(); }
void cassie_sim_set_dof_damping(cassie_sim_t *c, double *damp) { for (size_t i = 0; i < c->m_damp.size(); i++) c->m_damp[i] = damp[i]; }
void cassie_sim_set_body_mass(cassie_sim_t *c, double *mass) { for (size_t i = 0; i < c->m_mass.size(); i++) c->m_mass[i] = mass[i]; }
void cassie_sim_set_body_ipos(cassie_sim_t *c, double *ipos) {   // the reference indexes its argument with [i + j], not [3 i + j] (:1386-1393); kept
  const int nb = (int)c->m_mass.size(); for (int i = 0; i < nb; i++) for (int j = 0; j < 3; j++) c->m_ipos[3 * i + j] = ipos[i + j]; }
void cassie_sim_set_geom_friction(cassie_sim_t *c, double *fric) { for (size_t i = 0; i < c->m_fric.size(); i++) c->m_fric[i] = fric[i]; }
void cassie_sim_set_body_name_mass(cassie_sim_t *c, const char *name, double mass) { const int id = c->b->impl->hm.body_id(name ? name : ""); if (id >= 0) c->m_mass[id] = mass; }
double cassie_sim_get_body_name_mass(cassie_sim_t *c, const char *name) { const int id = c->b->impl->hm.body_id(name ? name : ""); return id >= 0 ? c->m_mass[id] : 0.0; }
void cassie_sim_set_body_name_ipos(cassie_sim_t *c, const char *name, double *ipos) { const int id = c->b->impl->hm.body_id(name ? name : ""); if (id >= 0) for (int k = 0; k < 3; k++) c->m_ipos[3 * id + k] = ipos[k]; }
double *cassie_sim_get_body_name_ipos(cassie_sim_t *c, const char *name) { const int id = c->b->impl->hm.body_id(name ? name : ""); return id >= 0 ? &c->m_ipos[3 * id] : nullptr; }
int cassie_sim_get_joint_num_dof(cassie_sim_t *c, const char *name) { const HostModel &hm = c->b->impl->hm; const int j = hm.joint_id(name ? name : ""); if (j < 0) return 0; return hm.jnt_type[j] == JNT_FREE ? 6 : (hm.jnt_type[j] == JNT_BALL ? 3 : 1); }
void cassie_sim_set_dof_name_damping(cassie_sim_t *c, const char *name, double *damp) { const HostModel &hm = c->b->impl->hm; const int j = hm.joint_id(name ? name : ""); if (j < 0) return; const int nd = cassie_sim_get_joint_num_dof(c, name); for (int i = 0; i < nd; i++) c->m_damp[hm.jnt_dofadr[j] + i] = damp[i]; }
double *cassie_sim_get_dof_name_damping(cassie_sim_t *c, const char *name) { const HostModel &hm = c->b->impl->hm; const int j = hm.joint_id(name ? name : ""); return j >= 0 ? &c->m_damp[hm.jnt_dofadr[j]] : nullptr; }
// the reference copies 3 values to &geom_friction[geom_id] (:1417-1421, not 3 * geom_id); here the geom's own triple is addressed
void cassie_sim_set_geom_name_friction(cassie_sim_t *c, const char *name, double *fric) { const int g = c->b->impl->hm.geom_id(name ? name : ""); if (g >= 0) for (int k = 0; k < 3; k++) c->m_fric[3 * g + k] = fric[k]; }
double *cassie_sim_get_geom_name_friction(cassie_sim_t *c, const char *name) { const int g = c->b->impl->hm.geom_id(name ? name : ""); return g >= 0 ? &c->m_fric[3 * g] : nullptr; }
void cassie_sim_just_set_const(cassie_sim_t *c) { sim_push(c); cassie_batch_set_const(c->b, nullptr, 0); }
void cassie_sim_set_const(cassie_sim_t *c) { sim_push(c); cassie_batch_set_const(c->b, nullptr, 1); sim_pull(c); }
void cassie_sim_full_reset(cassie_sim_t *c) {
  // src/cassiemujoco.c:2008-2033: qpos <- 35 constants, zero qvel / ctrl / applied forces / qacc, zero the torque delay line.
  // It does NOT touch time, the encoder filters or cassie_out, and does not call mj_forward.
  static const double q[35] = {0, 0, 1.01, 1, 0, 0, 0, 0.0045, 0, 0.4973, 0.9785, -0.0164, 0.01787, -0.2049, -1.1997, 0, 1.4267, 0, -1.5244, 1.5244, -1.5968,
                               -0.0045, 0, 0.4973, 0.9786, 0.00386, -0.01524, -0.2051, -1.1997, 0, 1.4267, 0, -1.5244, 1.5244, -1.5968};
  std::vector<double> cst(CST_W); c->b->impl->get("cst", cst.data());
  for (int i = 0; i < 60; i++) cst[CS_DELAY + i] = 0;
  c->b->impl->set("cst", cst.data());
  memcpy(c->qpos, q, sizeof q); memset(c->qvel, 0, sizeof c->qvel);   // 35 entries only (:2025): qpos beyond them (the cup of cassie_tray_box.xml) keeps its values
  c->b->impl->set("qpos", c->qpos); c->b->impl->set("qvel", c->qvel); cassie_batch_clear_forces(c->b);
  c->b->impl->reset_estimator(nullptr); c->b->impl->reset_estimator_device(nullptr);   // state_output_setup (src/cassiemujoco.c:2032) restarts the estimator's filters
  sim_pull(c);
}
// sizes the reference's Python wrapper asks for (src/cassiemujoco.c:1038-1060)
int cassie_sim_nbody(const cassie_sim_t *c) { return c->b->impl->hm.nbody; }
int cassie_sim_ngeom(const cassie_sim_t *c) { return c->b->impl->hm.ngeom; }
int cassie_sim_njnt(const cassie_sim_t *c) { return c->b->impl->hm.njnt; }
int cassie_sim_nu(const cassie_sim_t *c) { return c->b->impl->hm.nu; }
// src/cassiemujoco.c:1183-1189: qpos += h * qvel on the joint manifold; the reference then feeds an uninitialised cassie_out_t to its estimator,
// so *y carries no information there -- it is zeroed here
void cassie_integrate_pos(cassie_sim_t *c, state_out_t *y) { sim_push(c); cassie_batch_integrate_pos(c->b); c->b->impl->sync(); sim_pull(c); if (y) memset(y, 0, sizeof *y); }

// ---- geom placement (src/cassiemujoco.c:1466-1541; example/test_terrain.c:118-157 moves the stair boxes with these): borrowed pointers into host
// mirrors in the reference's geom numbering; the constant block is rebuilt before the next launch.  As in MuJoCo a new size leaves geom_rbound alone.
double *cassie_sim_geom_pos(cassie_sim_t *c) { return c->m_gpos.data(); }
double *cassie_sim_geom_quat(cassie_sim_t *c) { return c->m_gquat.data(); }
double *cassie_sim_geom_size(cassie_sim_t *c) { return c->m_gsize.data(); }
double *cassie_sim_geom_name_pos(cassie_sim_t *c, const char *name) { const int g = c->b->impl->hm.geom_id(name ? name : ""); return g >= 0 ? &c->m_gpos[3 * g] : nullptr; }
double *cassie_sim_geom_name_quat(cassie_sim_t *c, const char *name) { const int g = c->b->impl->hm.geom_id(name ? name : ""); return g >= 0 ? &c->m_gquat[4 * g] : nullptr; }
double *cassie_sim_geom_name_size(cassie_sim_t *c, const char *name) { const int g = c->b->impl->hm.geom_id(name ? name : ""); return g >= 0 ? &c->m_gsize[3 * g] : nullptr; }
void cassie_sim_set_geom_pos(cassie_sim_t *c, double *pos) { for (size_t i = 0; i < c->m_gpos.size(); i++) c->m_gpos[i] = pos[i]; }
void cassie_sim_set_geom_quat(cassie_sim_t *c, double *quat) { for (size_t i = 0; i < c->m_gquat.size(); i++) c->m_gquat[i] = quat[i]; }
void cassie_sim_set_geom_size(cassie_sim_t *c, double *size) { for (size_t i = 0; i < c->m_gsize.size(); i++) c->m_gsize[i] = size[i]; }
void cassie_sim_set_geom_name_pos(cassie_sim_t *c, const char *name, double *pos) { double *p = cassie_sim_geom_name_pos(c, name); if (p) for (int k = 0; k < 3; k++) p[k] = pos[k]; }
void cassie_sim_set_geom_name_quat(cassie_sim_t *c, const char *name, double *quat) { double *p = cassie_sim_geom_name_quat(c, name); if (p) for (int k = 0; k < 4; k++) p[k] = quat[k]; }
void cassie_sim_set_geom_name_size(cassie_sim_t *c, const char *name, double *size) { double *p = cassie_sim_geom_name_size(c, name); if (p) for (int k = 0; k < 3; k++) p[k] = size[k]; }
// the same for a batch: one placement shared by all its environments (pos / quat / size: NULL leaves that part alone); 0 / -1 (unknown geom)
int cassie_batch_set_geom_pose(cassie_batch_t *b, const char *name, const double *pos, const double *quat, const double *size) {
  HostModel &hm = b->impl->hm; const int g = hm.geom_id(name ? name : "");
  if (g < 0) { set_err("cassie_batch_set_geom_pose: unknown geom"); return -1; }
  if (pos) for (int k = 0; k < 3; k++) hm.geom_pos[3 * g + k] = pos[k];
  if (quat) for (int k = 0; k < 4; k++) hm.geom_quat[4 * g + k] = quat[k];
  if (size) for (int k = 0; k < 3; k++) hm.geom_size[3 * g + k] = size[k];
  return b->impl->rebuild_model() ? 0 : -1;
}

// ---- legacy verbs RL wrappers call around cassie_sim_step_pd (src/cassiemujoco.c:1072-1093, 1137-1145, 1159-1181, 1196-1225, 1974-2000, 2086-2092, 3380-3452)
// the dynamic subset of cassie_out_t from an observation row, the rest as cassie_out_init leaves it (:672-734)
static void fill_cassie_out(const double *o, const double *radio, cassie_out_t *y) {
  memset(y, 0, sizeof *y);
  y->isCalibrated = true;
  y->pelvis.medullaCounter = 1; y->pelvis.medullaCpuLoad = 159; y->pelvis.vtmTemperature = 40;
  y->pelvis.targetPc.etherCatStatus[1] = 8; y->pelvis.targetPc.etherCatStatus[4] = 1; y->pelvis.targetPc.taskExecutionTime = 2e-4; y->pelvis.targetPc.cpuTemperature = 60;
  y->pelvis.battery.dataGood = true; y->pelvis.battery.stateOfCharge = 1;
  for (int i = 0; i < 4; i++) y->pelvis.battery.temperature[i] = 30;
  for (int i = 0; i < 12; i++) y->pelvis.battery.voltage[i] = 4.2;
  y->pelvis.radio.radioReceiverSignalGood = true; y->pelvis.radio.receiverMedullaSignalGood = true;
  for (int i = 0; i < 16; i++) y->pelvis.radio.channel[i] = radio[i];
  y->pelvis.vectorNav.dataGood = true; y->pelvis.vectorNav.pressure = 101.325; y->pelvis.vectorNav.temperature = 25;
  for (int i = 0; i < 4; i++) y->pelvis.vectorNav.orientation[i] = o[OB_QUAT + i];
  for (int i = 0; i < 3; i++) { y->pelvis.vectorNav.angularVelocity[i] = o[OB_GYRO + i]; y->pelvis.vectorNav.linearAcceleration[i] = o[OB_ACCEL + i]; y->pelvis.vectorNav.magneticField[i] = o[OB_MAG + i]; }
  static const double tl[5] = {140.63, 140.63, 216.16, 216.16, 45.14}, gr[5] = {25, 25, 16, 16, 50};
  for (int sd = 0; sd < 2; sd++) {
    cassie_leg_out_t *leg = sd ? &y->rightLeg : &y->leftLeg; leg->medullaCounter = 1; leg->medullaCpuLoad = 94;
    elmo_out_t *dr[5] = {&leg->hipRollDrive, &leg->hipYawDrive, &leg->hipPitchDrive, &leg->kneeDrive, &leg->footDrive};
    for (int i = 0; i < 5; i++) { dr[i]->statusWord = 0x0637; dr[i]->dcLinkVoltage = 48; dr[i]->driveTemperature = 30; dr[i]->torqueLimit = tl[i]; dr[i]->gearRatio = gr[i];
      dr[i]->position = o[OB_MPOS + 5 * sd + i]; dr[i]->velocity = o[OB_MVEL + 5 * sd + i]; dr[i]->torque = o[OB_MTORQUE + 5 * sd + i]; }
    cassie_joint_out_t *jn[3] = {&leg->shinJoint, &leg->tarsusJoint, &leg->footJoint};
    for (int i = 0; i < 3; i++) { jn[i]->position = o[OB_JPOS + 3 * sd + i]; jn[i]->velocity = o[OB_JVEL + 3 * sd + i]; }
  }
}
// :2090-2092.  The observation row holds what `*y = c->cassie_out` copied out in the last step (:1127); before the first step it is the initial bus state
cassie_out_t cassie_sim_get_cassie_out(cassie_sim_t *c) { double o[OBS_W]; cassie_out_t y; c->b->impl->get("obs", o); fill_cassie_out(o, c->b->radio.data(), &y); return y; }
// :1137-1145: user torques -> safety layer -> motor model -> physics; y = the bus as the sensors filled it in this tick.  A pd_in_t that carries only
// motorPd.torque makes pd_input_step the identity on the torques, so the launch is the same one cassie_sim_step_pd uses
void cassie_sim_step(cassie_sim_t *c, cassie_out_t *y, const cassie_user_in_t *u) {
  pd_in_t pd; memset(&pd, 0, sizeof pd);
  for (int i = 0; i < 5; i++) { pd.leftLeg.motorPd.torque[i] = u->torque[i]; pd.rightLeg.motorPd.torque[i] = u->torque[5 + i]; }
  state_out_t so; sim_push(c); cassie_sim_step_pd_batch(c->b, &pd, &so); sim_pull(c);
  if (y) *y = cassie_sim_get_cassie_out(c);
}
// :1159-1181: the same tick with ONE physics step whatever the model's timestep is
void cassie_sim_step_pd_no2khz(cassie_sim_t *c, state_out_t *y, const pd_in_t *u) { sim_push(c); c->b->impl->nsub_override = 1; cassie_sim_step_pd_batch(c->b, u, y); c->b->impl->nsub_override = 0; sim_pull(c); }
// :1196-1204: the borrowed pointer is a host mirror, uploaded (constant block rebuilt) before the next launch
double *cassie_sim_timestep(cassie_sim_t *c) { return &c->timestep; }
void cassie_sim_set_timestep(cassie_sim_t *c, double dt) { c->timestep = dt; }
int cassie_batch_set_timestep(cassie_batch_t *b, double dt) { if (!(dt > 0)) return -1; b->impl->hm.timestep = dt; return b->impl->rebuild_model() ? 0 : -1; }
// :1221-1225
int cassie_sim_forward(cassie_sim_t *c) { sim_push(c); c->b->impl->step(0, 1); sim_pull(c); return 0; }
// :1974-2000: pin / free the pelvis with a stiff spring-damper on its three slides and damping on its ball joint
static void sim_hold(cassie_sim_t *c, bool on) {
  HostModel &hm = c->b->impl->hm; sim_push(c);
  for (int i = 0; i < 3 && i < hm.njnt; i++) { hm.jnt_stiffness[i] = on ? 1e5 : 0; if (on) hm.qpos_spring[i] = c->qpos[i]; }
  for (int i = 0; i < 6 && i < hm.nv; i++) { hm.dof_damping[i] = on ? (i < 3 ? 1e4 : 1e4) : 0; c->m_damp[i] = hm.dof_damping[i]; }
  c->b->impl->rebuild_model(); sim_push(c);
}
void cassie_sim_hold(cassie_sim_t *c) { sim_hold(c, true); }
void cassie_sim_release(cassie_sim_t *c) { sim_hold(c, false); }
// :3380-3452
cassie_state_t *cassie_state_alloc(void) { cassie_state *s = new cassie_state(); memset(s->qpos, 0, sizeof s->qpos); memset(s->qvel, 0, sizeof s->qvel); memset(s->qpos_snap, 0, sizeof s->qpos_snap); memset(s->qvel_snap, 0, sizeof s->qvel_snap); return s; }
void cassie_state_free(cassie_state_t *s) { if (!s) return; snap_delete(s->snap); delete s; }
double *cassie_state_time(cassie_state_t *s) { return &s->time; }
double *cassie_state_qpos(cassie_state_t *s) { return s->qpos; }
double *cassie_state_qvel(cassie_state_t *s) { return s->qvel; }
void cassie_get_state(const cassie_sim_t *cc, cassie_state_t *s) {
  cassie_sim_t *c = const_cast<cassie_sim_t *>(cc); BatchBase *impl = c->b->impl; sim_push(c);
  if (s->snap && !(s->snap->n == 1 && s->snap->qw == cassie_batch_row_width(c->b, "qpos") && s->snap->vw == cassie_batch_row_width(c->b, "qvel") && s->snap->device == impl->device)) { snap_delete(s->snap); s->snap = nullptr; }
  if (!s->snap) { s->snap = impl->snap_alloc(); if (!s->snap) return; }
  impl->snap_get(s->snap);
  memcpy(s->qpos, c->qpos, sizeof s->qpos); memcpy(s->qvel, c->qvel, sizeof s->qvel); s->time = c->time;
  memcpy(s->qpos_snap, s->qpos, sizeof s->qpos); memcpy(s->qvel_snap, s->qvel, sizeof s->qvel); s->time_snap = s->time;
}
void cassie_set_state(cassie_sim_t *c, const cassie_state_t *s) {
  if (!s->snap) { set_err("cassie_set_state: the state object holds no snapshot (cassie_get_state first)"); return; }
  if (!c->b->impl->snap_set(s->snap, nullptr)) return;   // any simulator of the same model accepts it
  sim_pull(c);
  // what the caller wrote through cassie_state_qpos / qvel / time after the snapshot was taken
  if (memcmp(s->qpos, s->qpos_snap, sizeof s->qpos)) memcpy(c->qpos, s->qpos, sizeof s->qpos);
  if (memcmp(s->qvel, s->qvel_snap, sizeof s->qvel)) memcpy(c->qvel, s->qvel, sizeof s->qvel);
  if (s->time != s->time_snap) c->time = s->time;
  sim_push(c);
}
void cassie_state_copy(cassie_state_t *dst, const cassie_state_t *src) {
  if (!src->snap) return;
  if (dst->snap && (dst->snap->n != src->snap->n || dst->snap->qw != src->snap->qw || dst->snap->vw != src->snap->vw || dst->snap->esz != src->snap->esz)) { snap_delete(dst->snap); dst->snap = nullptr; }
  if (!dst->snap) { dst->snap = snap_new(src->snap->device, src->snap->n, src->snap->esz, src->snap->qw, src->snap->vw); if (!dst->snap) return; }
  if (!snap_clone_into(dst->snap, src->snap)) { set_err("cassie_state_copy failed"); return; }
  memcpy(dst->qpos, src->qpos, sizeof dst->qpos); memcpy(dst->qvel, src->qvel, sizeof dst->qvel); dst->time = src->time;
  memcpy(dst->qpos_snap, src->qpos_snap, sizeof dst->qpos_snap); memcpy(dst->qvel_snap, src->qvel_snap, sizeof dst->qvel_snap); dst->time_snap = src->time_snap;
}
cassie_state_t *cassie_state_duplicate(const cassie_state_t *src) { cassie_state_t *s = cassie_state_alloc(); cassie_state_copy(s, src); return s; }
// :1072-1093: model (per-environment constants, height field, timestep) and full dynamic state
void cassie_sim_copy(cassie_sim_t *dst, const cassie_sim_t *csrc) {
  cassie_sim_t *src = const_cast<cassie_sim_t *>(csrc); sim_push(src); sim_push(dst);
  if (!dst->b->impl->copy_model_from(src->b->impl)) return;
  dst->m_mass = src->m_mass; dst->m_ipos = src->m_ipos; dst->m_damp = src->m_damp; dst->m_fric = src->m_fric;
  dst->m_mass_dev = dst->m_mass; dst->m_ipos_dev = dst->m_ipos; dst->m_damp_dev = dst->m_damp; dst->m_fric_dev = dst->m_fric;
  dst->m_gpos = dst->m_gpos_dev = src->m_gpos; dst->m_gquat = dst->m_gquat_dev = src->m_gquat; dst->m_gsize = dst->m_gsize_dev = src->m_gsize;
  dst->hfield = src->hfield; dst->hfield_dev = src->hfield_dev; dst->timestep = dst->timestep_dev = src->timestep;
  for (int i = 0; i < 16; i++) dst->b->radio[i] = src->b->radio[i];
  cassie_state_t *s = cassie_state_alloc(); cassie_get_state(src, s); cassie_set_state(dst, s); cassie_state_free(s);
}
cassie_sim_t *cassie_sim_duplicate(const cassie_sim_t *src) {
  cassie_sim_t *c = cassie_sim_init(src->path.c_str(), true);
  if (c) cassie_sim_copy(c, src);
  return c;
}
// batched snapshots: an opaque handle per batch (cassie_batch_state_free before cassie_batch_free)
void *cassie_batch_state_alloc(cassie_batch_t *b) { return b->impl->snap_alloc(); }
void cassie_batch_state_free(cassie_batch_t *b, void *state) { (void)b; snap_delete((Snap *)state); }
int cassie_batch_get_state(cassie_batch_t *b, void *state) { return state && b->impl->snap_get((Snap *)state) ? 0 : -1; }
int cassie_batch_set_state(cassie_batch_t *b, const void *state, const unsigned char *mask) { return state && b->impl->snap_set((const Snap *)state, mask) ? 0 : -1; }
#include "legacy_stubs.inc"
}  // extern "C"
