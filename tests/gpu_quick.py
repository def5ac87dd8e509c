"""quick GPU sanity + timing used during development (not a pytest file)"""
import importlib, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from conftest import PD_TARGET, PD_PGAIN, PD_DGAIN
P = importlib.import_module('cassie-mujoco-sim_b200')
import torch
for prec, n in ((P.FP32, 4096), (P.FP32, 16384), (P.FP64, 1024)):
    b = P.CassieBatch(n, precision=prec)
    b.set_pd(P.pd_rows(n, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    b.step(50); b.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(4): b.step(50)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c = b.counters()
    print('prec', prec, 'n', n, 'ms per 50 ticks %.3f' % (ms / 4), 'env-steps/s %.3e' % (n * 200 / (ms * 1e-3)), 'mean nefc %.1f iters %.1f' % (c[:, 0].mean(), c[:, 3].mean()), flush=True)
