#!/usr/bin/env python3
"""bench.py -- Cassie env-steps/s for the cassie_sim_step_pd hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--envs E]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one control tick (one cassie_sim_step_pd, 0.5 ms of simulated time, 1 physics sub-step) for EVERY environment of the
batch, i.e. one launch of the fused step kernel.  Workload at N=1 = BASELINE config 2: 4096 envs, cassie.xml flat floor, fixed
motor-PD targets, initial pelvis height/yaw jitter U(-0.01, 0.01) seed 0.  N>1: the same 4096 envs on
every GPU (weak scaling, environments are independent; no data-path collective).

value     kernel-only throughput: PD rows and state resident in HBM; inputs larger than L2 (independent copies of the batch stepped
          round-robin, > 1.5 x L2 in total), one contiguous CUDA-event region of K steps on the launching stream, max over ranks.  The K
          steps run as K / T launches of T <= 50 ticks (cassie_batch_step(b, T), PD rows held); single_tick_launches = T = 1.
          (l2_memset_flush_mode: the same kernel on one copy with a 256 MiB memset between launches, per-launch events.)
e2e       the same metric through the reference-shaped C-ABI call cassie_sim_step_pd_batch(envs, pd_in_t[] host, state_out_t[] host):
          host->device copy of every env's PD input and device->host read of every env's observation inside the timed region.
roofline  for the dominant kernel (cassie_step_kernel<float>), algorithmic bytes = persistent state in + out per env-step.
cpu_baseline / --impl reference: the CPU restatement in oracle/ (physics restated from MuJoCo 2.1.0 semantics + the reference's real
          closed Agility blocks when oracle/_ref/liboracle_ref.so exists) on all host cores.  The reference's own libcassiemujoco.so
          cannot be built: MuJoCo 2.1.0 is not available.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))

PD_TARGET = [0.0045, 0, 0.4973, -1.1997, -1.5968, -0.0045, 0, 0.4973, -1.1997, -1.5968]
PD_PGAIN = [70, 70, 100, 100, 50] * 2
PD_DGAIN = [7, 7, 8, 8, 5] * 2
STATE_BYTES_FP32 = 4 * ((36 + 32 + 32 + 192 + 52 + 8 + 96) + (36 + 32 + 32 + 192 + 96 + 96))   # rows read + rows written per env per launch
WORKLOAD = 'config2: 4096 envs/GPU cassie.xml flat floor, fixed motor-PD targets, pelvis z/yaw jitter U(-0.01,0.01) seed 0'


def jittered_qpos(q0, n, seed=0):
    rng = np.random.default_rng(seed)
    q = np.repeat(q0[None, :], n, axis=0).copy()
    q[:, 2] += rng.uniform(-0.01, 0.01, n)
    yaw = rng.uniform(-0.01, 0.01, n)
    q[:, 3] = np.cos(yaw / 2); q[:, 4] = 0; q[:, 5] = 0; q[:, 6] = np.sin(yaw / 2)
    return q


REF_TICKS = 250   # control ticks per reference-arm step


# ------------------------------------------------------------------------------ CPU arm (oracle)
def cpu_arm(n_threads, envs_per_thread, ticks, warm_ticks=0):
    """aggregate env-steps/s of the CPU oracle with one thread per core, each owning private sims."""
    import oracle as O
    ref = os.path.exists(O.lib_path(ref=True))
    L = O.load(ref=ref)
    model = os.path.join(REPO, 'tests', 'golden', 'cassie.omodel')
    u = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    groups = []
    for _ in range(n_threads):
        sims = (C.c_void_p * envs_per_thread)(*[L.osim_new(model.encode()) for _ in range(envs_per_thread)])
        groups.append(sims)

    def run(sims, t):
        L.osim_run(sims, envs_per_thread, C.byref(u), t)

    def run_all(t):
        th = [threading.Thread(target=run, args=(g, t)) for g in groups]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0
    if warm_ticks:
        run_all(warm_ticks)
    dt = run_all(ticks)
    for g in groups:
        for s in g:
            L.osim_free(s)
    return n_threads * envs_per_thread * ticks / dt, dt, ref


def reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # one private sim per host thread (the reference's own threading model; more sims per thread only thrash the caches), one bench
    # "step" = REF_TICKS control ticks of all of them
    ept, ticks = 1, REF_TICKS * args.steps
    val, dt, ref = cpu_arm(cores, ept, ticks, warm_ticks=REF_TICKS * min(args.warmup, 4))
    kind = 'port'
    sample = '%d envs (%d threads x %d private sim) x %d ticks (%d per step) of %s; physics = oracle/cassie_oracle.c (fp64 restatement of MuJoCo 2.1.0 semantics), Agility blocks = %s' % (
        cores * ept, cores, ept, ticks, REF_TICKS, WORKLOAD, 'the reference archive libagilitycassie.a incl. state_output_step' if ref else 'oracle twins')
    line = {'impl': 'reference', 'metric': 'Cassie env-steps/s', 'value': val, 'unit': 'env-steps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * dt / args.steps,   # one step = REF_TICKS ticks of `cores` environments 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'parallelism': 'cpu x%d threads' % cores},
            'cpu_baseline': {'value': val, 'unit': 'env-steps/s', 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': val, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line, default=lambda o: o.tolist() if hasattr(o, 'tolist') else str(o)), flush=True)


# ------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
                                          'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ''
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons)}


def gpu_arm(args, rank, local_rank, world):
    import torch
    P = importlib.import_module('cassie-mujoco-sim_b200')
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    n = args.envs
    b = P.CassieBatch(n, device=local_rank, precision=P.FP32)
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    q0 = b.qpos()[0]
    b.set_qpos(jittered_qpos(q0, n, seed=rank))
    b.forward()
    rows = P.pd_rows(n, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    b.set_pd(rows)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks['hbm_gbs'], 'measured (MEASURED_PEAKS.json)') if 'hbm_gbs' in peaks else (6650.0, 'fallback (B200_PROFILING.md)')

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only.  Inputs larger than L2: NB independent copies of the workload (different jitter seeds) are stepped round-robin, so
    # every launch finds its state rows in HBM, not in L2 (per copy ~3.5 KB/env; NB copies > 1.5 x the 126 MB L2).  One contiguous timed
    # region of K launches (K steps of the 4096-env batch), CUDA events on the launching stream.
    per_copy = n * (36 + 32 + 32 + 192 + 52 + 8 + 96 + 320 + 96 + 8) * 4
    nb_copies = max(2, -(-int(1.5 * 126e6) // per_copy))
    copies = [b]
    for k in range(1, nb_copies):
        bk = P.CassieBatch(n, device=local_rank, precision=P.FP32)
        bk.set_stream(torch.cuda.current_stream().cuda_stream)
        bk.set_qpos(jittered_qpos(q0, n, seed=1000 * k + rank)); bk.forward(); bk.set_pd(rows)
        copies.append(bk)
    # untimed set-up: every copy lands and settles into standing under the PD controller (600 ticks = 0.3 s of simulated time), so that the
    # timed region measures the steady workload (12 equality + 8 contact-pyramid rows, ~10 PGS sweeps) whatever K and W are
    for bk in copies:
        bk.step(600)
    for _ in range(args.warmup):
        for bk in copies:
            bk.step(1)
    launches0 = sum(bk.launch_count() for bk in copies)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    # K steps (control ticks of the whole batch) = K / T launches of T ticks each with the PD rows held (cassie_batch_step(b, T); config 2's
    # targets are constant).  T = the largest divisor of K not above 50 (50 ticks = one 40 Hz policy step of the reference's demos).
    T = max(t for t in range(1, 51) if args.steps % t == 0)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(args.steps // T):
        copies[i % nb_copies].step(T)
    k1.record()
    barrier()
    launches = sum(bk.launch_count() for bk in copies) - launches0
    ms_kernel = k0.elapsed_time(k1)
    # ---- the same K steps as K single-tick launches (reported beside the value)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(args.steps):
        copies[i % nb_copies].step(1)
    s1.record()
    barrier()
    ms_single = s0.elapsed_time(s1) / args.steps
    # ---- the same on ONE copy with a 256 MiB memset between launches (the other flush method; per-launch events; reported beside the value)
    nfl = min(args.steps, 50)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nfl)]
    for e0, e1 in ev:
        flush.zero_()
        e0.record(); b.step(1); e1.record()
    barrier()
    ms_flush = sum(e0.elapsed_time(e1) for e0, e1 in ev) / nfl
    for bk in copies[1:]:
        bk.close()
    # ---- end to end through the AoS C-ABI: pd_in_t[n] host -> step -> state_out_t[n] host, every step
    pd = (P.pd_in_t * n)()
    for e in range(n):
        for i in range(5):
            for leg, off in ((pd[e].leftLeg, 0), (pd[e].rightLeg, 5)):
                leg.motorPd.pTarget[i] = PD_TARGET[off + i]; leg.motorPd.pGain[i] = PD_PGAIN[i]; leg.motorPd.dGain[i] = PD_DGAIN[i]
    out = (P.state_out_t * n)()
    obs_t = b.torch_view('obs') if dist else None
    gathered = torch.empty((world * n, P.OBS_WIDTH), dtype=torch.float32, device='cuda') if dist else None
    e2e_steps = max(10, min(args.steps, 100))
    for _ in range(3):
        b.L.cassie_sim_step_pd_batch(b.h, C.byref(pd), C.byref(out))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        b.L.cassie_sim_step_pd_batch(b.h, C.byref(pd), C.byref(out))
    barrier()
    t_e2e = time.perf_counter() - t0
    # the one optional collective of the path: all-gather of every rank's fp32 observation block, timed on its own
    ms_gather = None
    if dist:
        for _ in range(3):
            dist.all_gather_into_tensor(gathered, obs_t)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(20):
            dist.all_gather_into_tensor(gathered, obs_t)
        g1.record(); barrier()
        ms_gather = g0.elapsed_time(g1) / 20
    clocks = sampler.stop() if sampler else None
    # ---- the HBM-bound integrate kernel (cassie_batch_integrate_pos) on a state larger than L2
    integ = None
    if rank == 0:
        nb = 1 << 20
        bi = P.CassieBatch(nb, device=local_rank, precision=P.FP32)
        bi.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            bi.integrate_pos()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            bi.integrate_pos()
        e1.record(); torch.cuda.synchronize()
        ms_i = e0.elapsed_time(e1) / reps
        bytes_i = nb * 4 * (35 + 32 + 35)
        integ = {'kernel': 'cassie_integrate_kernel<float>', 'bound': 'hbm', 'achieved': bytes_i / (ms_i * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                 'frac': bytes_i / (ms_i * 1e-3) / 1e9 / hbm_peak, 'traffic': 3.857e8,   # profiles/r1_integrate_kernel_ncu_summary.md
                 'envs': nb, 'bytes_per_env': 4 * (35 + 32 + 35), 'ms': ms_i}
        bi.close()
    # ---- the other BASELINE configs (parity-test cases, not the bench line): per-GPU slices, kernel-only, short runs
    others = None
    if rank == 0 and not args.no_extra:
        others = {}

        def timed(bb, nsteps, pre=None):
            bb.set_stream(torch.cuda.current_stream().cuda_stream)
            bb.step(300); bb.sync()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for i in range(nsteps):
                if pre:
                    pre(i)
                bb.step(1)
            a1.record(); torch.cuda.synchronize()
            c = bb.counters()
            return {'env_steps_per_s': bb.n * nsteps / (a0.elapsed_time(a1) * 1e-3), 'ms_per_tick': a0.elapsed_time(a1) / nsteps,
                    'mean_rows': float(c[:, 0].mean()), 'mean_pgs_sweeps': float(c[:, 3].mean()), 'dropped_contacts': int(c[:, 4].sum())}
        try:
            n3 = 16384
            b3 = P.CassieBatch(n3, device=local_rank, precision=P.FP32)
            b3.set_pd(P.pd_rows(n3, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
            rng = np.random.default_rng(1234)
            push = np.zeros((n3, 6)); push[:, :2] = rng.uniform(-100, 100, (n3, 2))

            def pushes(i):          # every 400 ticks: U(-100,100) N xy push on the pelvis held for 100 ticks, scaled to the short run
                if i % 40 == 0:
                    b3.apply_force(push, 'cassie-pelvis')
                elif i % 40 == 10:
                    b3.clear_forces()
            others['config3_16384_envs_pelvis_pushes'] = timed(b3, 80, pushes); b3.close()
            n4 = 8192
            b4 = P.CassieBatch(n4, modelfile=P.model_path('cassie_hfield'), device=local_rank, precision=P.FP32)
            terr = (np.random.default_rng(7).uniform(0, 1, (64, 200, 200)) * 0.25).astype(np.float32); terr[:, 95:105, 95:105] = 0
            b4.set_hfield_data(terr); b4.set_pd(P.pd_rows(n4, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
            others['config4_8192_envs_per_gpu_hfield_64_terrains_amp0.05m'] = timed(b4, 80); b4.close()
            b5 = P.CassieBatch(n4, modelfile=P.model_path('cassie_tray_box'), device=local_rank, precision=P.FP32)
            ph = np.random.default_rng(99).uniform(0, 2 * np.pi, (n4, 1)); amp = np.array([0.05, 0.05, 0.3, 0.4, 0.3] * 2) * 0.2
            rows5 = P.pd_rows(n4, pTarget=np.array(PD_TARGET) + amp * np.sin(ph + np.array([0] * 5 + [np.pi] * 5)), pGain=PD_PGAIN, dGain=PD_DGAIN)
            b5.set_pd(rows5)
            others['config5_8192_envs_per_gpu_tray_box'] = timed(b5, 80); b5.close()
            # SURVEY 8f-2 / 8f-3 on the config-2 workload: derived-quantity rows on, then per-env randomised constants + set_const (extended kernel instance)
            b6 = P.CassieBatch(n, device=local_rank, precision=P.FP32)
            b6.set_pd(P.pd_rows(n, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)); b6.set_qpos(jittered_qpos(b6.qpos()[0], n)); b6.forward()
            b6.enable_aux()
            others['config2_with_derived_quantity_rows'] = timed(b6, 80)
            rng6 = np.random.default_rng(5)
            b6.set_model('body_mass', b6.get_model('body_mass') * rng6.uniform(0.8, 1.2, (n, 1)))
            b6.set_model('dof_damping', b6.get_model('dof_damping') * rng6.uniform(0.5, 2.0, (n, 32)))
            fr = b6.get_model('geom_friction'); fr[:, 0::3] *= rng6.uniform(0.6, 1.1, (n, 1)); b6.set_model('geom_friction', fr)
            t0 = time.time(); b6.set_const(reset_state=True); b6.sync(); t_sc = time.time() - t0
            b6.set_qpos(jittered_qpos(b6.qpos()[0], n)); b6.forward()
            others['config2_with_randomised_constants_and_rows'] = dict(timed(b6, 80), set_const_ms_for_all_envs=1e3 * t_sc); b6.close()
        except Exception as ex:
            others['error'] = repr(ex)
    # ---- reduce over ranks
    t = torch.tensor([ms_kernel, t_e2e], dtype=torch.float64, device='cuda')
    per_rank = None
    if dist:
        allt = torch.empty((world, 2), dtype=torch.float64, device='cuda')
        dist.all_gather_into_tensor(allt, t)
        per_rank = {'kernel_ms_per_step': [float(x) / args.steps for x in allt[:, 0]], 'e2e_ms_per_step': [1e3 * float(x) / e2e_steps for x in allt[:, 1]]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_kernel, t_e2e = float(t[0]), float(t[1])
    if rank == 0:
        cores = os.cpu_count() or 1
        try:
            cpu_val, cpu_dt, ref = cpu_arm(cores, 1, 30000, warm_ticks=500)
            cpu = {'value': cpu_val, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
                   'sample': '%d envs (%d threads x 1 private sim) x 30000 ticks of the same workload (~1-2 s per core); oracle/cassie_oracle.c fp64 + %s' % (
                       cores, cores, 'reference Agility archive' if ref else 'Agility twins')}
        except Exception as ex:   # the oracle is a checker; its absence must not void the GPU number
            cpu = {'value': None, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': 'unavailable: %r' % (ex,)}
        ms_step = ms_kernel / args.steps
        value = world * n * args.steps / (ms_kernel * 1e-3)
        ach = STATE_BYTES_FP32 * n / (T * ms_step * 1e-3) / 1e9   # algorithmic bytes of one launch (state rows in + out, once per launch) / its duration
        line = {'metric': 'Cassie env-steps/s', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': WORKLOAD, 'envs_per_gpu': n, 'ticks_per_step': 1, 'ticks_per_launch': T, 'parallelism': 'env-sharded x%d (no data-path collective)' % world,
                           'l2': 'inputs larger than L2: %d independent copies of the batch (%.0f MB) stepped round-robin, one contiguous timed region' % (nb_copies, nb_copies * per_copy / 1e6),
                           'model': 'compiled table of model/cassie.xml'},
                'l2_memset_flush_mode': {'ms_per_step': ms_flush, 'env_steps_per_s_this_rank': n / (ms_flush * 1e-3), 'note': 'one copy, 256 MiB memset between launches, per-launch events'},
                'clocks': clocks,
                'e2e': {'value': world * n * e2e_steps / t_e2e, 'unit': 'env-steps/s', 'h2d_bytes_per_step': n * P.PD_WIDTH * 4, 'd2h_bytes_per_step': n * P.OBS_WIDTH * 4,
                        'api': 'cassie_sim_step_pd_batch(envs, pd_in_t[n] host, state_out_t[n] host), %d steps, host AoS pack/unpack included%s' % (
                            e2e_steps, ''), 'obs_allgather_ms': ms_gather},
                'gpu_launches': launches, 'per_rank': per_rank,
                'roofline': {'kernel': 'cassie_step_kernel<float>', 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak, 'traffic': 9.29e6,   # dram__bytes_read.sum + write.sum of a single-tick launch, profiles/r1_step_kernel_v6_ncu_summary.md
                             'peak_source': peak_src, 'bytes_per_env_per_launch': STATE_BYTES_FP32, 'ticks_per_launch': T,
                             'note': 'latency/issue-bound by design (SURVEY 8d): algorithmic HBM traffic is only the persistent state rows in+out'},
                'single_tick_launches': {'ms_per_step': ms_single, 'env_steps_per_s_this_rank': n / (ms_single * 1e-3), 'note': 'the same K steps as K launches of one tick'},
                'roofline_integrate': integ, 'cpu_baseline': cpu, 'other_configs': others}
        print(json.dumps(line, default=lambda o: o.tolist() if hasattr(o, 'tolist') else str(o)), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--envs', type=int, default=4096)
    ap.add_argument('--no-extra', dest='no_extra', action='store_true', help='skip the short runs of BASELINE configs 3-5')
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        reference_arm(args, rank, world)
    else:
        # host pack / unpack threads of the AoS entry point: share the host's cores between the ranks of this node (set before any
        # OpenMP runtime is loaded); the roofline traffic figure cites the ncu capture under profiles/
        os.environ.setdefault('CASSIE_B200_AOS_THREADS', str(max(2, min(32, (os.cpu_count() or 8) // (2 * max(1, world))))))
        gpu_arm(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
