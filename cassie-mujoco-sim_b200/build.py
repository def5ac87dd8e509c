"""Builds libcassie_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and by the Python host module."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libcassie_b200.so')
SRCS = [os.path.join(CSRC, 'cassie_b200.cu'), os.path.join(CSRC, 'mjcf.cpp')]
DEPS = SRCS + [os.path.join(CSRC, f) for f in ('step_core.inl', 'devmodel.h', 'devbuild.h', 'model.h', 'estimator_host.h', 'legacy_stubs.inc')] + [
    os.path.join(HERE, '..', 'include', 'cassie_b200.h'), os.path.join(HERE, '..', 'include', 'cassie_bus.h')]


def nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    return 'nvcc'


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return LIB
    cmd = [nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-shared', '-Xcompiler', '-fPIC,-fopenmp', '-lgomp',
           '-Xptxas', '-v' if verbose else '-O3', '-o', LIB] + SRCS
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stderr[-4000:])
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose='-v' in sys.argv))
