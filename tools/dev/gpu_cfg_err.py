"""distribution of the fp32 kernel's error against the fp64 oracle over the sampled environments of a BASELINE-size batch (dev aid for tests/test_gpu_configs.py)"""
import importlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests')); sys.path.insert(0, os.path.join(REPO, 'oracle'))
import bench, oracle as O
import test_gpu_configs as T
P = importlib.import_module('cassie-mujoco-sim_b200')
cfg = int(sys.argv[1]); H = int(sys.argv[2]) if len(sys.argv) > 2 else 600; every = 50
n = bench.CONFIGS[cfg]['envs']
W = bench.Workload(P, cfg, n, 0, 0, None); b = W.make_copy(0); q_init = b.qpos()
gait = None
if cfg == 5:
    f, ph = bench.philox_uniform(99, 0, n, 0.5, 1.5), bench.philox_uniform(99, 1, n, 0.0, 2 * np.pi); gait = (f, ph[:, None] + np.array([0.0] * 5 + [np.pi] * 5)[None, :])
sample = np.sort(np.random.default_rng(100 + cfg).choice(n, 64, replace=False))
got = []
for t in range(H):
    W.before_step(b, 0, 1); b.step(1)
    if (t + 1) % every == 0: got.append(b.qpos()[sample])
got = np.array(got)
err = np.zeros((got.shape[0], 64)); arg = np.zeros((got.shape[0], 64), dtype=int)
for k, e in enumerate(sample):
    want = T.oracle_replay(O, cfg, W, q_init[e], int(e), H, every, gait)
    err[:, k] = np.abs(got[:, k, :] - want).max(axis=1); arg[:, k] = np.abs(got[:, k, :] - want).argmax(axis=1)
for i in range(err.shape[0]):
    r = err[i]; print('cfg %d tick %4d: median %.1e  p90 %.1e  max %.1e (env %d)' % (cfg, (i + 1) * every, np.median(r), np.percentile(r, 90), r.max(), sample[r.argmax()]) + '  qpos index %d' % arg[i, r.argmax()])
