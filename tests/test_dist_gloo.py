"""world_size-2 gloo test of the host-side multi-GPU logic: env sharding and the optional observation all-gather."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, REPO)
    import importlib
    P = importlib.import_module('cassie-mujoco-sim_b200')
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    start, count = P.env_shard(n_total, rank, world)
    local = torch.arange(start, start + count, dtype=torch.float32)[:, None] * torch.ones(1, P.OBS_WIDTH)
    allobs = P.gather_observations(local)
    ok = allobs.shape == (n_total, P.OBS_WIDTH) and torch.equal(allobs[:, 0], torch.arange(n_total, dtype=torch.float32))
    q.put((rank, bool(ok), start, count))
    dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [8, 9])
def test_shard_and_gather_world2(n_total):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_total) % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert sum(r[3] for r in res) == n_total


def test_aos_thread_team_leaves_headroom_in_the_cpu_quota():
    """N ranks x their AoS thread teams must stay below the job's CPU quota (the teams busy-wait; at the quota the whole cgroup is throttled: measured on
    the 8-GPU lease, 8 x 12 threads on 96 CPUs)"""
    import bench
    for quota, world, pinned in ((96, 8, 12), (48, 4, 12), (24, 2, 12), (16, 1, 128), (128, 8, 16), (4, 2, 2)):
        t = bench.aos_threads_for_rank(pinned, quota // world)
        assert 2 <= t <= 32 and t <= pinned
        if quota // world > 4:
            assert t * world <= quota - 2 * world
