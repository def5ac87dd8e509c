// step_kernel.cuh -- the fused step kernel (one warp per environment) and its argument block.  Included by step_inst.cu, which is compiled once per
// kernel instance (precision x plain / extended x model features) so that the instances build in parallel and each carries only the code its
// models need, and by cassie_b200.cu for the argument types; the instances are reached through cassie_step_entry_<tag>() function pointers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "devmodel.h"
#include "estimator_host.h"
#include "step_core.inl"

namespace cassie {

#ifdef CASSIE_DBG_ALL
constexpr bool DBG_IN_ALL_INSTANCES = true;
#else
constexpr bool DBG_IN_ALL_INSTANCES = false;
#endif

template <typename real> struct EnvArrays {
  real *qpos, *qvel, *qacc_ws, *cst, *pd, *xfrc, *obs, *dbg, *qM, *aux, *cenv, *task, *gait; const real *pd_host; real *obs_host; double *est; int *dfilt, *counters, *ticket; const float *hfield; const unsigned char *mask; int cta_sync, nsub, warp_stride, env0; int n, n_terrain, qpos_w, qvel_w, ystride, xb; size_t hfield_stride;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// stage `bytes` (multiple of 16) from global to shared with one TMA bulk copy; all threads of the CTA return after it landed
__device__ __forceinline__ void tma_stage(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  const uint32_t bar_a = smem_u32(bar), dst_a = smem_u32(dst_smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a), "l"(src_gmem), "r"(bytes), "r"(bar_a) : "memory");
  }
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(bar_a), "r"(0) : "memory");
}

template <typename real> __host__ __device__ constexpr size_t model_bytes() { return (sizeof(DevModel<real>) + 127) / 128 * 128; }
template <typename real> __host__ __device__ constexpr size_t warp_bytes(int ystride, bool ext) { return ((size_t)(ext ? scratch_reals_ext(ystride) : scratch_reals(ystride)) * sizeof(real) + 127) / 128 * 128; }

// mode 0: step nticks; mode 1: mj_forward only.  INST: 0 plain instance, 1 extended instance (per-env constants, derived quantities, task PD, set_const,
// estimator), 2 plain instance + estimator stage (what cassie_sim_step_pd_batch runs when nothing else asks for the extended one)
template <typename real, int INST, int FEAT>
__global__ void __launch_bounds__(512) cassie_step_kernel(const DevModel<real> *__restrict__ gmodel, EnvArrays<real> A, int nticks, int mode) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  DevModel<real> *cmp = reinterpret_cast<DevModel<real> *>(smem_raw);
  tma_stage(cmp, gmodel, (uint32_t)model_bytes<real>(), &bar);
  const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
  real *sm = reinterpret_cast<real *>(smem_raw + model_bytes<real>() + warp * A.warp_stride);   // warp_stride = warp_bytes(ystride, instance), precomputed on the host: cheap to rematerialise under register pressure
  const int qw = (FEAT & F_XB) ? A.qpos_w : QPOS_W_MAIN, vw = (FEAT & F_XB) ? A.qvel_w : QVEL_W_MAIN;
  const DevModel<real> &cm = *cmp;
  // persistent CTAs with a static round-robin schedule: in round r CTA b owns environments [(r gridDim + b) nwarps, ... + nwarps), one per warp.  The
  // grid is sized to the resident CTAs, the environments cost about the same, and a static schedule lets every warp pull the rows of its NEXT
  // environment towards L2 while it steps the current one.  (Multi-tick launches: the CTA's warps walk the stages together, STAGE_SYNC, so that
  // they share instruction-cache lines -- the kernel is ~200 KB of code, several times the L1.5.)
  const int nwarps = blockDim.x >> 5, sync_on = (mode == 0 && (nticks > 1 || (A.cta_sync & 32))) ? (A.cta_sync & 31) : 0;   // bit 5: also in single-tick launches
  const int stride = gridDim.x * nwarps;
  for (int base = blockIdx.x * nwarps; base < A.n; base += stride) {
    const int env = A.env0 + base + warp;   // a launch covers environments [env0, env0 + n)
    if (!(A.cta_sync & 64)) __syncthreads();   // the warps of a CTA start every round together: they then share instruction-cache lines through the round (measured: 16 384
                       // environments, 7 rounds, -20 % without it), and the multi-tick rendezvous counts below assume it
    { const int nenv = base + warp + stride < A.n ? env + stride : 0x7fffffff;   // warm L2 with the next round's state rows of this warp (about 2 KB per environment, one 128-byte line per lane)
      if (nenv != 0x7fffffff && mode == 0) {
        const char *p = nullptr;
        if (l < 6) p = (const char *)(A.cst + (size_t)nenv * CST_W) + 128 * l * (sizeof(real) / 4);
        else if (l < 9) p = (const char *)(A.dfilt + (size_t)nenv * DFILT_W) + 128 * (l - 6);
        else if (l < 11) p = (const char *)(A.pd + (size_t)nenv * PD_W) + 128 * (l - 9) * (sizeof(real) / 4);
        else if (l < 13) p = (const char *)(A.qpos + (size_t)nenv * qw) + 128 * (l - 11) * (sizeof(real) / 4);
        else if (l < 15) p = (const char *)(A.qvel + (size_t)nenv * vw) + 128 * (l - 13) * (sizeof(real) / 4);
        else if (l < 17) p = (const char *)(A.qacc_ws + (size_t)nenv * vw) + 128 * (l - 15) * (sizeof(real) / 4);
        else if (l == 17) p = (const char *)(A.xfrc + (size_t)nenv * XFRC_W);
        if (p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      } }
    const bool active = base + warp < A.n && !(A.mask && !A.mask[env]);   // masked launches (reset / set_const of a subset)
    if (!active) {   // keep the rendezvous count of the working warps
      if (sync_on) { const int per = __popc(sync_on); for (int i = 0; i < nticks * (A.nsub > 0 ? A.nsub : cm.nsub) * per; ++i) __syncthreads(); }
      continue;
    }
    // warm the L2/L1 path of the rows that are addressed in place later (controller state, PD row, FIR taps)
    if (l < 6) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.cst + (size_t)env * CST_W + 32 * l));
    else if (l < 9) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.dfilt + (size_t)env * DFILT_W + 32 * (l - 6)));
    else if (l < 11) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.pd + (size_t)env * PD_W + 32 * (l - 9)));
    else if (l == 11) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.xfrc + (size_t)env * XFRC_W));
    // optional zero-copy input: this environment's motor-PD row straight from mapped host memory into its device row (the AoS entry point
    // uses one DMA instead: 2 x 4096 small PCIe reads at kernel start were measured slower than the copy engine; kept for C2C-attached hosts)
    if (A.pd_host) { for (int i = l; i < PD_W; i += 32) A.pd[(size_t)env * PD_W + i] = A.pd_host[(size_t)env * PD_W + i]; }
    // state rows: qpos to shared memory, qvel / warm start one value per lane; everything else is addressed in place
    for (int i = l; i < qw; i += 32) sm[S_QPOS + i] = (mode == 3) ? cm.qpos0[i] : A.qpos[(size_t)env * qw + i];   // set_const works at the reference configuration
    real qvel = A.qvel[(size_t)env * vw + l], qacc_ws = A.qacc_ws[(size_t)env * vw + l], xqvel = 0, xqacc_ws = 0;
    if ((FEAT & F_XB) && A.xb >= 0 && l < 6) { xqvel = A.qvel[(size_t)env * vw + 32 + l]; xqacc_ws = A.qacc_ws[(size_t)env * vw + 32 + l]; }
    __syncwarp();
    EnvPtrs<real> E;   // (E.dbg: the stage dumps are compiled into the extended instance only; batches created with debug = 1 run that one)
    E.cst = A.cst + (size_t)env * CST_W; E.dfilt = A.dfilt + (size_t)env * DFILT_W; E.pd = A.pd + (size_t)env * PD_W; E.xfrc = A.xfrc + (size_t)env * XFRC_W; E.task = A.task ? A.task + (size_t)env * TASK_W : nullptr; E.gait = A.gait ? A.gait + (size_t)env * GAIT_W : nullptr;
    E.hfield = A.hfield ? A.hfield + (size_t)(env % A.n_terrain) * A.hfield_stride : nullptr;
    E.obs = A.obs + (size_t)env * OBS_W; E.qM = A.qM + (size_t)env * 2 * NM_MAX; E.dbg = (DBG_IN_ALL_INSTANCES || INST == 1) && A.dbg ? A.dbg + (size_t)env * D_SIZE : nullptr; E.counters = A.counters + (size_t)env * 8;
    E.aux = A.aux ? A.aux + (size_t)env * AUX_W : nullptr; E.cenv = A.cenv ? A.cenv + (size_t)env * CE_W : nullptr; E.cta_sync = sync_on; E.nsub = A.nsub;
    E.est = A.est ? A.est + (size_t)env * EST_W : nullptr; E.est_out = E.obs + OB_EST_OUT;
    step_env<real, INST == 1, FEAT, INST >= 1>(cm, sm, E, qvel, qacc_ws, xqvel, xqacc_ws, nticks, mode);
    __syncwarp();
    if (mode >= 2) continue;   // query / set_const: only the aux row / the constant row is written
    for (int i = l; i < qw; i += 32) A.qpos[(size_t)env * qw + i] = sm[S_QPOS + i];
    A.qvel[(size_t)env * vw + l] = qvel; A.qacc_ws[(size_t)env * vw + l] = qacc_ws;
    if ((FEAT & F_XB) && A.xb >= 0 && l < 6) { A.qvel[(size_t)env * vw + 32 + l] = xqvel; A.qacc_ws[(size_t)env * vw + 32 + l] = xqacc_ws; }
    // ... and its observation row goes back the same way as soon as the environment is done, overlapped with the environments still stepping (sending it
    // right after the controller stage, before the physics, was measured: no end-to-end gain, -2.8 % on every instance from the longer-lived pointer)
    if (A.obs_host) { for (int i = l; i < OBS_W; i += 32) A.obs_host[(size_t)env * OBS_W + i] = A.obs[(size_t)env * OBS_W + i]; }
    __syncwarp();
  }
}

}  // namespace cassie
