"""BASELINE C5 as specified: E0 in {5, 10, 15, 20, 25, 30} keV x {Al, Si, Cu, Au} x {continuous Joy-Luo, continuous
explicit Bethe, DS}, 1e7 electrons each, at 1 / 2 / 4 / 8 GPUs (strong scaling: the 1e7 ids of a cell are split over
the ranks, ONE all-reduce of the histogram||counter block per cell).

    python scripts/run_c5.py [--el 10000000] > profiles/<tag>_c5_1gpu.jsonl
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/run_c5.py

Stated parameters (BASELINE.json leaves them open): tilt 70 deg, Emin = max(500 eV, 0.05 E0) for continuous runs and
E0 / 2 for DS, maxScatt = 10000, default table tolerances (tol_E 1e-4, tol_W 1e-7), histograms + counters only.
One JSON line per cell: wall time incl. the table build, electrons/s, events/s, backscatter coefficient, SHA-1 of the
reduced statistics block (equal at every N)."""
import argparse, hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                           # noqa: E402
import torch.distributed as dist                       # noqa: E402
from mcdiscrete_b200 import _lib                       # noqa: E402
from mcdiscrete_b200.engine import Engine              # noqa: E402
from mcdiscrete_b200.probTables import genTables       # noqa: E402
from mcdiscrete_b200.distributed import allreduce_stats, rank_range   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--el', type=int, default=10 ** 7)
    ap.add_argument('--materials', default='Al,Si,Cu,Au')
    a = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('LOCAL_RANK', 0), ('WORLD_SIZE', 1)))
    torch.cuda.set_device(local)
    if world > 1:
        from mcdiscrete_b200.distributed import bind_to_gpu
        bind_to_gpu(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tmax(x):
        t = torch.tensor([float(x)], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    t_all = time.time()
    digests = []
    for mat in a.materials.split(','):
        for E0 in (5000., 10000., 15000., 20000., 25000., 30000.):
            for mode, bm in (('cont', 'JL'), ('cont', 'explicit'), ('DS', 'classical')):
                emin = max(500., 0.05 * E0) if mode == 'cont' else E0 / 2
                ip = dict(num_el=a.el, E0=E0, Emin=emin, material=mat, s_tilt=70., thickness=3e6, elastic='Ruth_vanilla',
                          mode=mode, Bethe_model=bm, diffMFP=False, gen_tables=True, tol_E=1e-4, tol_W=1e-7, maxScatt=10000,
                          Wc=10., electron_output=[], scatter_output=[], seed=12345)
                sync()
                t0 = time.time()
                eng = Engine(ip, device=local)
                if mode == 'DS':
                    genTables(ip, engine=eng)
                torch.cuda.synchronize()
                t1 = time.time()
                lo, n = rank_range(a.el, rank, world)
                for _ in eng.run_batches(n, 1_000_000, first_id=lo, accumulate=True, els=False):
                    pass
                tot = eng.pipeline_stats()
                allreduce_stats(tot)
                blk = tot.cpu().numpy()
                sync()
                t2 = time.time()
                wall, run = tmax(t2 - t0), tmax(t2 - t1)
                c = blk[_lib.HIST_TOTAL:]
                assert int(c[0]) == a.el
                sha = hashlib.sha1(blk.tobytes()).hexdigest()[:16]
                digests.append(sha)
                if rank == 0:
                    print(json.dumps(dict(config='C5', n_gpus=world, material=mat, E0=E0, Emin=emin, mode=mode,
                                          bethe=bm if mode == 'cont' else None, electrons=a.el, wall_s=wall, run_s=run,
                                          tables_s=t1 - t0, electrons_per_s=a.el / run, events_per_s=int(c[1]) / run,
                                          events_per_electron=int(c[1]) / a.el, eta=int(c[3]) / a.el,
                                          long_frac=int(c[6]) / a.el, nan_clamped=int(c[10]), stats_sha1=sha,
                                          kernel=eng.launch_info())), flush=True)
                eng.close()
    if rank == 0:
        print(json.dumps(dict(config='C5 summary', n_gpus=world, cells=len(digests), electrons_per_cell=a.el,
                              total_wall_s=time.time() - t_all,
                              sha1_of_digests=hashlib.sha1(''.join(digests).encode()).hexdigest()[:16])), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
