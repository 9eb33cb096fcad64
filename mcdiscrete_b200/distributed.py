"""Multi-GPU: one process per GPU (torchrun), electrons partitioned by id range, disjoint Philox streams
by construction (counter = electron id), and ONE collective: the all-reduce of the int64
histogram||counter block.  Raw els / scats rows stay rank-local."""
import os


def rank_range(num_el, rank, world):
    """Electron ids of `rank`: [r*N/R, (r+1)*N/R) -- contiguous, disjoint, covering [0, N)."""
    lo = (num_el * rank) // world
    hi = (num_el * (rank + 1)) // world
    return lo, hi - lo


def init_from_env(backend=None):
    """torch.distributed from torchrun's environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def allreduce_stats(stats):
    """Sum the int64 histogram||counter block over ranks, in place (NCCL over NVLink on GPUs)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def run_partitioned(engine, num_el, first_id=0, **kw):
    """Simulate this rank's share of electrons [first_id, first_id+num_el) and all-reduce the statistics.
    Returns (RunResult of the local share, global counters dict, global histograms dict)."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, n = rank_range(num_el, rank, world)
    res = engine.run(count=n, first_id=first_id + lo, **kw)
    stats = engine.stats_tensor()
    allreduce_stats(stats)
    return res, engine.counters(), engine.histograms()
