"""Host-side logic of the multi-GPU path on CPU: id-range partition and the statistics all-reduce over
gloo with world_size 2 (the GPU path runs the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcdiscrete_b200.distributed import rank_range, allreduce_stats


def test_rank_range_partitions_exactly():
    for n in (1, 7, 1000, 10 ** 8, 10 ** 8 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [rank_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(c for _, c in spans) == n
            for (lo, c), (lo2, _) in zip(spans, spans[1:]):
                assert lo + c == lo2
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mcdiscrete_b200 import _lib
    n = _lib.HIST_TOTAL + 23
    rng = np.random.default_rng(rank)
    local = torch.from_numpy(rng.integers(0, 10 ** 9, n).astype(np.int64))
    mine = local.clone()
    allreduce_stats(local)
    lo, cnt = rank_range(1001, rank, world)
    # a job with fewer electrons than ranks: the rank whose id range is empty still joins the run's ONE all-reduce
    # with a zero block (doScatter.main starts from zeros; a rank that skipped the collective would hang the others)
    lo1, cnt1 = rank_range(1, rank, world)
    one = torch.zeros(n, dtype=torch.int64)
    one[_lib.HIST_TOTAL] += cnt1                     # electrons_done of this rank's share
    allreduce_stats(one)
    assert int(one[_lib.HIST_TOTAL]) == 1
    q.put((rank, mine.numpy(), local.numpy(), lo, cnt))
    dist.destroy_process_group()


def test_stats_allreduce_world2_gloo():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    total = got[0][1] + got[1][1]
    assert np.array_equal(got[0][2], total) and np.array_equal(got[1][2], total)
    assert got[0][3] == 0 and got[0][4] + got[1][4] == 1001 and got[1][3] == got[0][4]
