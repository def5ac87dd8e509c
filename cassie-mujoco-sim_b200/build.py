"""Builds libcassie_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and by the Python host module.

The fused step kernel is compiled once per instance (precision x plain / extended x model features; csrc/step_inst.cu with -D flags), the
instances and the two host-side sources in parallel, then linked into one shared object."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, '_obj')
LIB = os.path.join(HERE, 'libcassie_b200.so')
DEPS = [os.path.join(CSRC, f) for f in ('cassie_b200.cu', 'mjcf.cpp', 'step_inst.cu', 'step_kernel.cuh', 'step_core.inl', 'cassie_tree_gen.inc', 'devmodel.h', 'devbuild.h', 'model.h',
                                        'estimator_host.h', 'legacy_stubs.inc')] + [
    os.path.join(HERE, '..', 'include', 'cassie_b200.h'), os.path.join(HERE, '..', 'include', 'cassie_bus.h')]
# (tag, real, instance: 0 plain / 1 extended / 2 plain + estimator, feature set): features 1 = extra free body, 2 = height field, 4 = box geoms (csrc/devmodel.h F_*)
INSTANCES = [('f00', 'float', 0, 0), ('f10', 'float', 1, 0), ('f02', 'float', 0, 2), ('f12', 'float', 1, 2), ('f04', 'float', 0, 4), ('f14', 'float', 1, 4), ('f05', 'float', 0, 5), ('f15', 'float', 1, 5),
             ('f07', 'float', 0, 7), ('f17', 'float', 1, 7), ('d07', 'double', 0, 7), ('d17', 'double', 1, 7),
             ('f20', 'float', 2, 0), ('f22', 'float', 2, 2), ('f24', 'float', 2, 4), ('f25', 'float', 2, 5), ('f27', 'float', 2, 7), ('d27', 'double', 2, 7)]
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC,-fopenmp'] + os.environ.get('CASSIE_B200_NVCC_EXTRA', '').split()   # extra flags for A/B builds


def nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    return 'nvcc'


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed: %s\n%s' % (' '.join(cmd), r.stderr[-6000:]))
    return r.stderr if verbose else ''


def build(force=False, verbose=False, out=None):
    lib = out or LIB
    if not out and not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    ptx = ['-Xptxas', '-v'] if verbose else []
    jobs = []
    for tag, real, dr, feat in INSTANCES:
        jobs.append([nvcc()] + ARCH + ptx + ['-DINST_REAL=' + real, '-DINST_DR=%d' % dr, '-DINST_FEAT=%d' % feat, '-DINST_TAG=' + tag, '-c', os.path.join(CSRC, 'step_inst.cu'),
                                           '-o', os.path.join(OBJ, 'step_%s.o' % tag)])
    jobs.append([nvcc()] + ARCH + ptx + ['-c', os.path.join(CSRC, 'cassie_b200.cu'), '-o', os.path.join(OBJ, 'cassie_b200.o')])
    jobs.append([nvcc()] + ARCH + ['-c', os.path.join(CSRC, 'mjcf.cpp'), '-o', os.path.join(OBJ, 'mjcf.o')])
    with ThreadPoolExecutor(max_workers=max(2, min(len(jobs), (os.cpu_count() or 4)))) as ex:
        logs = list(ex.map(lambda c: _run(c, verbose), jobs))
    if verbose:
        sys.stderr.write(''.join(logs))
    objs = [j[-1] for j in jobs]
    _run([nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-Xcompiler', '-fPIC,-fopenmp', '-lgomp', '-o', lib + '.tmp'] + objs, verbose)
    os.replace(lib + '.tmp', lib)   # atomic: a reader never sees a half-written library
    return lib


if __name__ == '__main__':
    print(build(force=True, verbose='-v' in sys.argv, out=sys.argv[sys.argv.index('--out') + 1] if '--out' in sys.argv else None))
