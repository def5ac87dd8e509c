"""quick GPU sanity + timing used during development (not a pytest file)"""
import importlib, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from conftest import PD_TARGET, PD_PGAIN, PD_DGAIN
P = importlib.import_module('cassie-mujoco-sim_b200')
import torch
cfgs = [(P.FP32, 4096, 1), (P.FP32, 4096, 50), (P.FP32, 16384, 10), (P.FP64, 1024, 10)]
if len(sys.argv) > 1:
    cfgs = [(P.FP32, 4096, 1), (P.FP32, 4096, 50)]
for prec, n, nt in cfgs:
    b = P.CassieBatch(n, precision=prec)
    rng = np.random.default_rng(0)
    jit = 0.05 if os.environ.get('JITTER') else 0.0
    b.set_pd(P.pd_rows(n, pTarget=np.array(PD_TARGET) + rng.uniform(-jit, jit, (n, 10)), pGain=PD_PGAIN, dGain=PD_DGAIN))
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    if os.environ.get('AUX'): b.enable_aux()
    for _ in range(3): b.step(nt)
    b.sync()
    reps = max(4, 200 // nt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): b.step(nt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c = b.counters()
    print('WPB', os.environ.get('CASSIE_B200_WPB', 'auto'), 'prec', prec, 'n', n, 'ticks/launch', nt, 'ms/tick %.4f' % (ms / reps / nt), 'env-steps/s %.3e' % (n * reps * nt / (ms * 1e-3)),
          'mean nefc %.1f iters %.1f' % (c[:, 0].mean(), c[:, 3].mean()), flush=True)
