"""GPU edge cases: ragged batch sizes, the full BASELINE sizes through size-independent properties, fault containment, bad arguments."""
import ctypes as C

import numpy as np
import pytest

from conftest import PD_DGAIN, PD_PGAIN, PD_TARGET, product

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def P():
    return product()


def _run(P, n, ticks, per_launch=1, precision=None, jitter=None):
    b = P.CassieBatch(n, precision=P.FP32 if precision is None else precision)
    if jitter is not None:
        q = b.qpos()
        q[:, 2] += jitter
        b.set_qpos(q)
        b.forward()
    b.set_pd(P.pd_rows(n, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
    for _ in range(ticks // per_launch):
        b.step(per_launch)
    return b


def test_result_does_not_depend_on_batch_size_or_slot(P):
    """an environment's trajectory is a function of its own state only: n = 1, 17, one more than the resident slots, 4096"""
    ref = _run(P, 1, 60).qpos()[0]
    for n in (17, 148 * 16 + 1, 4096):
        q = _run(P, n, 60).qpos()
        assert np.array_equal(q[0], ref) and np.array_equal(q[-1], ref), n
    assert np.array_equal(_run(P, 5, 60, per_launch=20).qpos()[3], ref)


def test_permutation_invariance_at_full_size(P):
    """BASELINE config 2 size: shuffling which slot holds which initial state only shuffles the results"""
    n, rng = 4096, np.random.default_rng(0)
    jit = rng.uniform(-0.01, 0.01, n)
    perm = rng.permutation(n)
    a, b = _run(P, n, 120, jitter=jit), _run(P, n, 120, jitter=jit[perm])
    qa, qb = a.qpos(), b.qpos()
    assert np.array_equal(qa[perm], qb)
    assert np.isfinite(qa).all() and np.unique(qa[:, 2]).size > n // 2
    c = a.counters()
    assert (c[:, 4] == 0).all() and (c[:, 0] <= 48).all()


def test_largest_baseline_size_runs(P):
    """65 536 environments on one device (BASELINE config 5's total): a few ticks, every copy identical and finite"""
    b = _run(P, 65536, 3)
    q = b.qpos()
    assert np.isfinite(q).all() and (q == q[0]).all()


def test_nan_in_one_environment_stays_there(P):
    n = 64
    b = P.CassieBatch(n, precision=P.FP32)
    b.set_pd(P.pd_rows(n, pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN))
    q = b.qpos()
    q[7, 9] = np.nan
    b.set_qpos(q)
    b.step(50)
    q = b.qpos()
    assert not np.isfinite(q[7]).all()
    ok = np.delete(q, 7, axis=0)
    assert np.isfinite(ok).all() and (ok == ok[0]).all()


def test_masked_reset_leaves_the_others_bit_identical(P):
    n = 6
    a, b = _run(P, n, 150), _run(P, n, 150)
    mask = np.zeros(n, dtype=np.uint8)
    mask[[1, 4]] = 1
    b.reset(mask)
    fresh = P.CassieBatch(1, precision=P.FP32)
    assert np.array_equal(b.qpos()[1], fresh.qpos()[0]) and b.time()[4] == 0
    a.step(40)
    b.step(40)
    keep = [0, 2, 3, 5]
    assert np.array_equal(a.qpos()[keep], b.qpos()[keep]) and np.array_equal(a.obs()[keep], b.obs()[keep])
    assert not np.array_equal(a.qpos()[1], b.qpos()[1])


def test_bad_arguments_fail_loudly(P):
    L = P.lib()
    assert not L.cassie_batch_init(b'/nonexistent/model.xml', 4, 0, 0) and b'model' in L.cassie_b200_last_error().lower()
    assert not L.cassie_batch_init(P.model_path().encode(), 0, 0, 0)
    assert not L.cassie_batch_init(P.model_path().encode(), 4, 99, 0) and b'device' in L.cassie_b200_last_error().lower()
    b = P.CassieBatch(3)
    before = b.qpos()
    assert L.cassie_batch_apply_force(b.h, (C.c_double * 18)(*([50.0] * 18)), b'no-such-body') == -1
    b.step(5)
    c = P.CassieBatch(3)
    c.step(5)
    assert np.array_equal(b.qpos(), c.qpos()) and not np.array_equal(before, b.qpos())
    with pytest.raises((RuntimeError, ValueError)):
        b.set_model('body_mass', np.ones((3, L.cassie_batch_nbody(b.h) + 1)))
