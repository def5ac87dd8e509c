"""workload for an ncu capture of the step kernel: BASELINE config C settled into its steady state, then single-tick launches.
usage (under ncu, -k regex:cassie_step_kernel -s <skip> -c 1):  python tools/dev/prof_step.py [config=2] [envs] ; EST=1 runs the in-kernel estimator"""
import importlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench
import torch
P = importlib.import_module('cassie-mujoco-sim_b200')
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg]['envs']
W = bench.Workload(P, cfg, n, 0, 0, torch)
b = W.make_copy(0)
if os.environ.get('EST'):
    b.enable_estimator_device(True)
if os.environ.get('AUX'):
    b.enable_aux()
b.step(600); b.sync()            # launch 2 (after the init forward launches)
for i in range(12):
    W.before_step(b, 0, 1); b.step(1)
b.sync()
c = b.counters()
print('config', cfg, 'envs', n, 'launches', b.launch_count(), 'mean rows %.1f sweeps %.1f' % (c[:, 0].mean(), c[:, 3].mean()))
