"""dev aid: step time vs warps per CTA for a few batch sizes (not a pytest file)"""
import importlib, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import numpy as np, torch
    sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
    from conftest import PD_TARGET, PD_PGAIN, PD_DGAIN
    P = importlib.import_module('cassie-mujoco-sim_b200')
    for n in (4096, 8192, 16384, 32768):
        b = P.CassieBatch(n)
        rng = np.random.default_rng(0)
        b.set_pd(P.pd_rows(n, pTarget=np.array(PD_TARGET) + rng.uniform(-0.05, 0.05, (n, 10)), pGain=PD_PGAIN, dGain=PD_DGAIN))
        b.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(300): b.step(1)
        b.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): b.step(1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 200
        print('wpb', os.environ.get('CASSIE_B200_WPB', 'auto'), 'n', n, 'ms/tick %.4f' % ms, 'Msteps/s %.2f' % (n / ms / 1e3), flush=True)
        b.close()
else:
    for w in (14, 15, 16):
        subprocess.run([sys.executable, __file__, 'child'], env=dict(os.environ, CASSIE_B200_WPB=str(w)))
