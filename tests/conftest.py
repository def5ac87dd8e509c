import importlib
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, 'oracle'), os.path.join(REPO, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, 'tests', 'golden')
REFERENCE = os.environ.get('CASSIE_REFERENCE', '/root/reference')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _have_gpu():
    try:
        import ctypes
        n = ctypes.c_int(0)
        for name in ('libcudart.so', 'libcudart.so.12', 'libcudart.so.13'):
            try:
                rt = ctypes.CDLL(name)
                return rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
            except OSError:
                continue
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # gpu-marked tests skip (instead of failing inside cassie_batch_init) on a machine without a CUDA device
    if any('gpu' in it.keywords for it in items) and not _have_gpu():
        skip = pytest.mark.skip(reason='no CUDA device on this machine (the stepper has no CPU fallback)')
        for it in items:
            if 'gpu' in it.keywords:
                it.add_marker(skip)


def product():
    """the product package (its directory name has a hyphen, so it is imported by string)."""
    return importlib.import_module('cassie-mujoco-sim_b200')


@pytest.fixture(scope='session')
def pkg():
    return product()


@pytest.fixture(scope='session')
def oracle_mod():
    import oracle as O
    O.build(ref=False)
    return O


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, 'model'))


# the BASELINE config-2 controller: fixed motor-PD targets / gains
PD_TARGET = [0.0045, 0, 0.4973, -1.1997, -1.5968, -0.0045, 0, 0.4973, -1.1997, -1.5968]
PD_PGAIN = [70, 70, 100, 100, 50] * 2
PD_DGAIN = [7, 7, 8, 8, 5] * 2
