"""N>1 path on CPU: two processes over gloo (127.0.0.1).  Each rank steps ITS shard of a global env range (device
replaced by the test-only host build of the step function) and the union must equal the unsharded run: the shard
map must not matter, and timings/counters must aggregate as max / sum over ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_agent_rl_for_fjsp_b200 import dist as fdist

TOTAL, SEED, STEPS = 21, 31337, 45


def _digest_env(g):
    from oracle.fjsp_oracle import philox_actions
    from tests.host_harness.hostharness import HostEnv

    e = HostEnv()
    e.reset(orders=None, num_orders=30, seed=SEED, genv=g, episode=0)
    rsum = 0.0
    for t in range(STEPS):
        _, _, r, _ = e.step(philox_actions(SEED, g, t))
        rsum += float(r.sum())
    return int(e.words().astype(np.uint64).sum()), rsum


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, lr, w = fdist.init(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = fdist.shard_range(TOTAL, rank, world)
    local = [_digest_env(g) for g in range(lo, hi)]
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, local))
    t_max = fdist.max_over_ranks(1.0 + rank)
    n_sum = fdist.sum_over_ranks(hi - lo)
    fdist.barrier()
    if rank == 0:
        q.put((gathered, t_max, n_sum))
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, t_max, n_sum = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t_max == 2.0 and n_sum == TOTAL
    ranges = [(lo, hi) for lo, hi, _ in gathered]
    assert ranges == [(0, 11), (11, 21)]
    union = [d for _, _, loc in gathered for d in loc]
    single = [_digest_env(g) for g in range(TOTAL)]
    assert union == single


def test_shard_range_properties():
    for total in (1, 7, 64, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 4, 8):
            parts = [fdist.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        fdist.shard_range(10, 2, 2)


def test_numa_binding_helper_is_safe_without_a_gpu():
    import os

    before = os.sched_getaffinity(0)
    cpus = fdist.bind_to_gpu_numa(0)  # no NVML device here: must return [] and leave the affinity alone
    assert isinstance(cpus, list)
    if not cpus:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
