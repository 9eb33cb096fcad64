"""Multi-GPU: one process per GPU (torchrun), electrons partitioned by id range, disjoint Philox streams
by construction (counter = electron id), and ONE collective: the all-reduce of the int64
histogram||counter block.  Raw els / scats rows stay rank-local."""
import os


def rank_range(num_el, rank, world):
    """Electron ids of `rank`: [r*N/R, (r+1)*N/R) -- contiguous, disjoint, covering [0, N)."""
    lo = (num_el * rank) // world
    hi = (num_el * (rank + 1)) // world
    return lo, hi - lo


def bind_to_gpu(local_rank):
    """Pin this process to the CPUs next to its GPU (NVML's ideal CPU set = the NUMA node of the GPU's PCIe root) so
    that the pinned host buffers of the drain are allocated on that node (first touch) and the D2H copies of the
    ranks do not all cross the socket interconnect: with 8 unbound ranks the aggregated D2H rate of one box was 90 GB/s
    (1.6x one link), profiles/rnd2m_bench_8gpu.json.  Returns a dict describing what was done; never raises
    (MCD_NO_BIND=1 or a restricted cpuset: nothing is changed).  `restore_affinity(info)` undoes it."""
    info = {'bound': False}
    if os.environ.get('MCD_NO_BIND'):
        info['why'] = 'MCD_NO_BIND'
        return info
    try:
        import pynvml
        import torch
        before = sorted(os.sched_getaffinity(0))
        info['before'] = len(before)
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(local_rank)
            bus = '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        want = sorted(cpus & set(before))
        if want and len(want) < len(before):
            os.sched_setaffinity(0, want)
            info.update(bound=True, cpus=len(want), first_cpu=want[0], last_cpu=want[-1], _restore=before)
        else:
            info['why'] = 'ideal CPU set covers the whole allowed set' if want else 'ideal CPUs are outside the cpuset'
    except Exception as ex:                                   # noqa: BLE001 -- binding is an optimisation only
        info['why'] = '%s: %s' % (type(ex).__name__, ex)
    return info


def restore_affinity(info):
    """Give the process its original CPU set back (e.g. before a host-side leg that should use every core)."""
    if info.get('_restore'):
        try:
            os.sched_setaffinity(0, info['_restore'])
        except Exception:                                     # noqa: BLE001
            pass


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
