"""GPU: the four GPU BASELINE configurations at their stated per-GPU sizes, fp32 throughput build against the fp64 oracle on 64 sampled environments.

The batches are built by bench.py's own Workload class (same seeds, jitter, terrains, push schedule, gaits as the benchmark lines), stepped in
single-tick launches with the host-driven events of the workload in between; every sampled environment is replayed by the oracle with the same
initial state, the same per-tick PD inputs (for config 5: the gait targets the kernel generates itself), pushes and terrain.

Tolerance: max |dqpos| <= 1e-4 over 600 control ticks from the initial drop for every configuration -- the north_star
tolerance; measured 3e-6 .. 9e-6 in the host emulation (tests/test_config_parity_emu.py).  Configs 4 and 5 use this repository's own height-field /
box contact rules: parity means kernel == oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO, product

sys.path.insert(0, REPO)
import bench  # noqa: E402

pytestmark = pytest.mark.gpu
HORIZON, NSAMPLE, TOL = 600, 64, 1e-4


def oracle_replay(O, cfg, W, q_init, e, horizon, check_every, gait=None):
    """qpos of environment e at ticks check_every, 2 check_every, ... under configuration cfg, from the fp64 oracle"""
    model = bench.CONFIGS[cfg]['model']
    o = O.OracleSim(os.path.join(GOLDEN, model + '.omodel'))
    nq, n = o.nq, W.n
    if W.terrains is not None:
        t = np.ascontiguousarray(W.terrains[e % W.terrains.shape[0]])
        C.memmove(o.L.osim_hfield_data(o.h), t.ctypes.data, t.nbytes)
    o.arr('qpos')[:nq] = q_init
    o.forward()
    u = O.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    out = []
    for t in range(horizon):
        if cfg == 3:     # bench.Workload.before_step with one copy: event k at tick 400 k, held 100 ticks
            if t % 400 == 0:
                o.arr('xfrc_applied').reshape(-1, 6)[1, :2] = bench.philox_uniform(1234, t // 400, (n, 2), -100.0, 100.0)[e]
            elif t % 400 == 100:
                o.arr('xfrc_applied')[:] = 0
        if cfg == 5:
            f, phase = gait
            u = O.make_pd(pTarget=np.array(PD_TARGET) + np.array(bench.GAIT_AMP) * np.sin(2 * np.pi * f[e] * t * 0.0005 + phase[e]), pGain=PD_PGAIN, dGain=PD_DGAIN)
        o.step_pd(u)
        if (t + 1) % check_every == 0:
            out.append(o.arr('qpos')[:nq].copy())
    return np.array(out)


@pytest.mark.parametrize('cfg', [2, 3, 4, 5])
def test_fp32_batch_of_baseline_size_follows_the_oracle(oracle_mod, cfg):
    P, O = product(), oracle_mod
    n = bench.CONFIGS[cfg]['envs']
    W = bench.Workload(P, cfg, n, 0, 0, None)
    b = W.make_copy(0)
    assert b.n == n and b.precision == P.FP32
    q_init = b.qpos()
    gait = None
    if cfg == 5:
        f, ph = bench.philox_uniform(99, 0, n, 0.5, 1.5), bench.philox_uniform(99, 1, n, 0.0, 2 * np.pi)
        gait = (f, ph[:, None] + np.array([0.0] * 5 + [np.pi] * 5)[None, :])
    sample = np.sort(np.random.default_rng(100 + cfg).choice(n, NSAMPLE, replace=False))
    every = 100
    got = []
    for t in range(HORIZON):
        W.before_step(b, 0, 1)
        b.step(1)
        if (t + 1) % every == 0:
            got.append(b.qpos()[sample])
    got = np.array(got)                                     # [checks, NSAMPLE, nq]
    worst, moved = 0.0, 0.0
    for k, e in enumerate(sample):
        want = oracle_replay(O, cfg, W, q_init[e], int(e), HORIZON, every, gait)
        worst = max(worst, float(np.abs(got[:, k, :] - want).max()))
        moved = max(moved, float(np.abs(want[-1] - q_init[e]).max()))
    print('config %d: %d envs, %d sampled, %d ticks: max|dqpos| fp32 kernel vs fp64 oracle = %.2e' % (cfg, n, NSAMPLE, HORIZON, worst))
    assert worst < TOL, (cfg, worst)
    assert moved > 0.05                                     # the robots really landed / were pushed / walked
    c = b.counters()
    assert int(c[:, 4].sum()) == 0                          # no contact was dropped for capacity anywhere in the batch
    if cfg == 3:     # the push of event 0 was really applied per environment (Philox keyed by env / event)
        assert np.abs(got[0, 0, :2] - got[0, 1, :2]).max() > 1e-4
    b.close()
