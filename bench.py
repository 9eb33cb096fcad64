#!/usr/bin/env python
"""bench.py -- headline benchmark of the per-electron trajectory loop (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload = "C2"): BASELINE.json configs[1] -- Si, E0 = 20 keV, DS mode
(Rutherford + Moller + Gryzinski + Quinn), default table tolerances (tol_E 1e-4, tol_W 1e-7), 70 deg tilt,
num_el = 1e6 per GPU; Emin = 10 keV and maxScatt = 1000 from the reference's inputFile/input.file (BASELINE.json
leaves them open).  One step = one pass of the trajectory kernel over 1e6 fresh electrons of the synthetic
pencil beam (ids never repeat across steps), histograms + counters fused, one els row per electron.
Weak scaling: every rank simulates its own 1e6 electrons (disjoint Philox streams); the only collective is the
int64 all-reduce of the histogram||counter block (NCCL).

Printed JSON: `value` = electrons/s with tables resident in HBM, device-timed (CUDA events, max over
ranks); `e2e` = the same through the public call Engine.run() with pinned HOST output buffers, i.e.
including per step the launch-parameter upload, the D2H copy of every els row, histograms and counters;
`roofline` for the trajectory kernel against the MEASURED FP32 / MUFU peaks of this GPU; `cpu_baseline` =
the oracle's C restatement of the reference loop on the host cores (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_EL = 1_000_000
WORKLOAD = dict(num_el=NUM_EL, E0=20000., Emin=10000., material='Si', s_tilt=70., thickness=3e6,
                elastic='Ruth_vanilla', mode='DS', Bethe_model='classical', diffMFP=False, gen_tables=True,
                tol_E=1e-4, tol_W=1e-7, maxScatt=1000, Wc=10.,
                electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'], scatter_output=[], seed=12345)
# algorithmic work per DS event: F = 181 + 24 S flops, M = 11 + 3 S MUFU ops, S = core shells
N_SHELLS = 3
FLOPS_PER_EVENT = 181 + 24 * N_SHELLS
MUFU_PER_EVENT = 11 + 3 * N_SHELLS


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every 5 ms from
    a thread (nvidia-smi -lms cannot deliver more than a sample or two inside a 100 ms region); nvidia-smi is the
    fallback when NVML is unavailable."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    BITS = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self._stop = index, [], None, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[self.index]) if vis and vis.split(',')[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    reasons = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((time.time(), mhz, reasons, pw))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(',')]))

    def stop(self, t0, t1):
        if self.nvml is not None:
            self._stop = True
            rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
            if not rows:
                return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['no samples']}
            sm = sorted(r[1] for r in rows)
            reasons = sorted(k for k, b in self.BITS.items() if any(r[2] & b for r in rows))
            return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': self.max_mhz, 'reasons': reasons, 'samples': len(rows),
                    'power_w_max': max(r[3] for r in rows), 'source': 'nvml, 5 ms polling inside the timed region'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(rows[0][2]), 'reasons': sorted(reasons),
                'samples': len(rows), 'power_w_max': max(float(r[3]) for r in rows), 'source': 'nvidia-smi -lms 20'}


def cpu_sample(n_el, nthreads, tables=None):
    """The oracle's C restatement of the reference loop on the host cores; returns (el/s, events/s, dt, tables)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import mcd_oracle as orc
    import mcd_oracle_cbind as oc
    if tables is None:
        tables = orc.gen_tables(dict(WORKLOAD, tol_E=1e-3, tol_W=1e-5))
    oc.run(WORKLOAD, 10 ** 9, 64, 1, tables, nthreads=nthreads)           # warm-up (library load, page-in)
    t0 = time.time()
    r = oc.run(WORKLOAD, 10 ** 9 + 1000, n_el, 1, tables, nthreads=nthreads)
    dt = time.time() - t0
    return n_el / dt, float(r['n_ev'].sum()) / dt, dt, tables


_PY_TABLES = None


def _py_job(first_id):
    """One job of the reference's pool: numTrajPerJob = 4 electrons (doScatter.py:216-220), Python port."""
    import mcd_oracle as orc
    els = orc.run_electrons(WORKLOAD, first_id, 4, 1, _PY_TABLES)
    return sum(len(e.scats) for e in els)


def python_pool_sample(tables, cores, jobs_per_core=60):
    """The Python restatement under multiprocessing.Pool(cores), one job = 4 electrons like the reference's
    MapScatterer (doScatter.py:22-95): the reference's own parallel structure at the reference's own speed class."""
    global _PY_TABLES
    import multiprocessing as mp
    _PY_TABLES = tables
    njobs = cores * jobs_per_core
    with mp.get_context('fork').Pool(cores) as pool:
        pool.map(_py_job, [7 * 10 ** 9 + 4 * j for j in range(cores)])             # warm-up: imports, page-in
        t0 = time.time()
        ev = sum(pool.map(_py_job, [6 * 10 ** 9 + 4 * j for j in range(njobs)], chunksize=1))
        dt = time.time() - t0
    return 4 * njobs / dt, ev / dt, dt, 4 * njobs


def run_reference(args, rank):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores.  The reference is
    pure Python and /root/reference does not exist on the GPU box, so this is the oracle port: the C restatement
    (all host threads) is the timed arm; the Python restatement, which mirrors the reference's per-event object
    structure and runs at the reference's own speed (~100-250 us/event/core), is reported beside it."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_el = 400000                      # ~2.5 s of work per step on 16 threads
    tables = None
    vals, evs = [], []
    for s in range(args.warmup + args.steps):
        el_s, ev_s, dt, tables = cpu_sample(n_el, cores, tables)
        if s >= args.warmup:
            vals.append(el_s)
            evs.append(ev_s)
    v = sum(vals) / len(vals)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import mcd_oracle as orc
    t0 = time.time()
    py = orc.run_electrons(WORKLOAD, 5 * 10 ** 9, 24, 1, tables)
    dtp = time.time() - t0
    try:
        pel, pev, pdt, pn = python_pool_sample(tables, cores)
        py_pool = {'value': pel, 'unit': 'electrons/s', 'events_per_s': pev, 'cores': cores,
                   'note': 'the same Python restatement in a multiprocessing pool of %d workers, 4 electrons per job like '
                           'the reference (doScatter.py:216-220): %d electrons in %.1f s' % (cores, pn, pdt)}
    except Exception as ex:          # noqa: BLE001
        py_pool = {'unavailable': str(ex)}
    sample = ('%d electrons per step of workload C2 (Si 20 keV DS tilt 70 Emin 10 keV), oracle C restatement, %d '
              'threads; tables at tol_E 1e-3 / tol_W 1e-5 (numpy oracle; lookup cost does not depend on table size)'
              % (n_el, cores))
    print(json.dumps({
        'impl': 'reference', 'metric': 'electron trajectories/s', 'value': v, 'unit': 'electrons/s', 'n_gpus': 0,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * n_el / v, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'events_per_s': sum(evs) / len(evs),
        'config': {'workload': 'C2', 'material': 'Si', 'E0_eV': 20000, 'Emin_eV': 10000, 'mode': 'DS',
                   's_tilt_deg': 70, 'num_el_per_step': n_el, 'maxScatt': 1000},
        'cpu_baseline': {'value': v, 'unit': 'electrons/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'python_port': {'value': 24 / dtp, 'unit': 'electrons/s', 'cores': 1,
                        'note': 'pure-Python restatement with the reference\'s per-event structure, 24 electrons'},
        'python_port_pool': py_pool,
        'e2e': {'value': v, 'unit': 'electrons/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--num-el', type=int, default=NUM_EL)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-drain', action='store_true', help='skip the output-drain leg (C3 shape, N = 1 only)')
    ap.add_argument('--drain-el', type=int, default=1_250_000, help='electrons per two batches of the output-drain leg')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        return run_reference(args, rank)

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: mcdiscrete_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from mcdiscrete_b200 import _lib
    from mcdiscrete_b200.engine import Engine
    from mcdiscrete_b200.probTables import genTables
    from mcdiscrete_b200.distributed import allreduce_stats, rank_range

    n_el = args.num_el
    ip = dict(WORKLOAD, num_el=n_el)
    eng = Engine(ip, device=local)
    t0 = time.time()
    tabs = genTables(ip, engine=eng)
    torch.cuda.synchronize()
    t_tables = time.time() - t0
    table_bytes = sum(4 * len(T.Es) * (len(T.Ws) - 1) for T in [tabs['Moller']] + tabs['Gryz'])
    # NPIPE contexts sharing the tables, one CUDA stream each: consecutive steps are pipelined, so the blocks of
    # step k+1 fill the SMs that the tail of step k leaves idle (1e6 electrons are only 5.9 per resident lane) and
    # the tail hand-over passes of one step (sparse: the last electrons of a step are serial chains of up to
    # ~870 events) run behind the main pass of another
    NPIPE = 6
    engs = [eng] + [eng.clone() for _ in range(NPIPE - 1)]
    tstreams = [torch.cuda.Stream() for _ in range(NPIPE)]
    for e in engs:
        e.set_capacity(n_el, 0)
    stream = torch.cuda.current_stream().cuda_stream
    stats = [e.stats_tensor() for e in engs]

    # measured compute roofs of this GPU (FMA and MUFU chains)
    import ctypes as C
    f32p, mufu, mhz = C.c_float(), C.c_float(), C.c_float()
    _lib.check(eng.lib.mcd_measure_peaks(local, C.byref(f32p), C.byref(mufu), C.byref(mhz)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_id = [0]

    def first_id():
        # fresh electrons every step, disjoint across ranks
        fid = (step_id[0] * world + rank) * n_el
        step_id[0] += 1
        return fid

    def device_step(k):
        e, ts = engs[k % NPIPE], tstreams[k % NPIPE]
        with torch.cuda.stream(ts):
            e.reset_stats(ts.cuda_stream)
            e.launch(first_id(), n_el, ts.cuda_stream)
            if world > 1:
                allreduce_stats(stats[k % NPIPE])

    # ---- device-timed throughput ------------------------------------------------------------------------
    for k in range(args.warmup):
        device_step(k)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    kernel_ms = []
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(NPIPE)]
    tw0 = time.time()
    ev0.record(tstreams[0])
    for ts in tstreams[1:]:
        ts.wait_event(ev0)                           # all streams start after the same timestamp
    for k in range(args.steps):
        device_step(k)
    for ts, e1 in zip(tstreams, ev1):
        e1.record(ts)
    barrier()
    tw1 = time.time()
    ms = max(ev0.elapsed_time(e1) for e1 in ev1)
    clocks = sampler.stop(tw0, tw1)
    tmax = torch.tensor([ms], device='cuda')
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    cnt = engs[(args.steps - 1) % NPIPE].counters()   # of the last step (summed over ranks when world > 1)
    events_per_step = cnt['events']           # whole job
    electrons_per_step = cnt['electrons_done']
    assert electrons_per_step == n_el * world, (electrons_per_step, n_el, world)
    value = electrons_per_step * args.steps / (ms * 1e-3)
    events_s = events_per_step * args.steps / (ms * 1e-3)

    # kernel-only duration (CUDA events around the trajectory kernel inside mcd_run, on its launch stream)
    for _ in range(args.steps):
        eng.reset_stats(stream)
        eng.launch(first_id(), n_el, stream)
        eng.sync(stream)
        kernel_ms.append(eng.kernel_ms())
    k_ms = sum(kernel_ms) / len(kernel_ms)
    launches = eng.launch_info()['launches_per_run']
    ev_local = eng.counters()['events']
    # roofline timing: the trajectory launches are the only kernels of a step, so the per-step device time of the
    # timed region above (steps overlapping on NPIPE streams, CUDA events on the launch streams) IS the average
    # duration of one step's launches under the conditions the throughput is quoted for; the serial stand-alone
    # duration of one step's launches (k_ms, includes its un-overlapped tail) is reported beside it
    step_ms = ms / args.steps
    ev_rate = ev_local / (step_ms * 1e-3)     # this rank
    fp32_ach = ev_rate * FLOPS_PER_EVENT / 1e12
    mufu_ach = ev_rate * MUFU_PER_EVENT / 1e12
    fp32_frac, mufu_frac = fp32_ach / f32p.value, mufu_ach / mufu.value
    traffic = None
    prof = os.path.join(ROOT, 'profiles', 'r1_traj_kernel_summary.json')
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get('dram_bytes_per_launch')
        except Exception:
            traffic = None
    bound = 'mufu' if mufu_frac >= fp32_frac else 'fp32'
    roofline = {'bound': bound,
                'achieved': mufu_ach if bound == 'mufu' else fp32_ach,
                'peak': mufu.value if bound == 'mufu' else f32p.value,
                'unit': 'Ttranscendental/s' if bound == 'mufu' else 'TFLOP/s',
                'frac': max(mufu_frac, fp32_frac), 'traffic': traffic,
                'kernel': 'traj_kernel_pool<3> (main pass + %d tail hand-over passes of the same kernel)' % (launches - 1),
                'kernel_ms': step_ms, 'kernel_ms_standalone': k_ms, 'events_per_launch': ev_local,
                'timing': 'per-step device time inside the pipelined timed region; standalone = one step launched alone',
                'fp32': {'achieved_tflops': fp32_ach, 'peak_tflops': f32p.value, 'frac': fp32_frac,
                         'flops_per_event': FLOPS_PER_EVENT},
                'mufu': {'achieved_tops': mufu_ach, 'peak_tops': mufu.value, 'frac': mufu_frac,
                         'ops_per_event': MUFU_PER_EVENT},
                'peak_source': 'measured on this GPU by mcd_measure_peaks (FFMA / MUFU.EX2 dependent chains); '
                               'MEASURED_PEAKS.json has no FP32/MUFU figure',
                'hbm': {'table_bytes_resident': table_bytes, 'peak_gbs_measured': 6547.2,
                        'achieved_gbs': (traffic / (step_ms * 1e-3) / 1e9) if traffic else None,
                        'frac': (traffic / (step_ms * 1e-3) / 1e9 / 6547.2) if traffic else None,
                        'note': 'DRAM traffic of the trajectory kernel (ncu) over its CUDA-event duration; the kernel is '
                                'issue-bound, HBM is second order'}}

    # ---- end to end through the public API with pinned host outputs ------------------------------------------
    h2d = C.sizeof(_lib.Params) + 1024      # launch parameter block: there are no per-step input arrays
    d2h = n_el * 32 + 8 * (_lib.HIST_TOTAL + C.sizeof(_lib.Counters) // 8)
    def e2e_pass(nsteps):
        last = None
        hook = (lambda e, ts: allreduce_stats(e.stats_tensor())) if world > 1 else None
        for r in eng.run_batches(nsteps * n_el, n_el, first_id=first_id() + 10 ** 12, on_launched=hook):
            last = r                                  # rows are in pinned host memory here
            _ = int(r.els['code'][0]) + r.counters['events']
        return last
    e2e_pass(NPIPE + 2)                               # warm every context of the pipeline (device buffers, pinned staging)
    barrier()
    t0 = time.time()
    r = e2e_pass(args.steps)
    barrier()
    dt = time.time() - t0
    tdt = torch.tensor([dt], device='cuda')
    if world > 1:
        dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
    e2e = n_el * world * args.steps / float(tdt.item())
    eta = float((r.outcome == 1).mean())
    cnt_e2e = r.counters

    # ---- output drain (north_star (3)): BASELINE C3 shape = Cu, full scatter_output (type, position, energy), rows
    # regrouped on the device into per-column arrays and drained to pinned host memory, pipelined over two contexts.
    # Reported beside the headline (N = 1 only): rows/s and the D2H GB/s the drain sustains. ---------------------------
    drain = None
    if world == 1 and not args.no_drain:
        ipc = dict(WORKLOAD, material='Cu', scatter_output=['type', 'position', 'energy'], num_el=args.drain_el)
        ec = Engine(ipc, device=local)
        genTables(ipc, engine=ec)
        nb = args.drain_el // 2
        for _ in ec.run_batches(2 * nb, nb, first_id=10 ** 13, ordered='columns'):      # warm-up: pin the staging buffers
            pass
        torch.cuda.synchronize()
        t0 = time.time()
        rows = 0
        for rr in ec.run_batches(4 * nb, nb, first_id=0, ordered='columns'):
            rows += len(rr.scats)
            _ = float(rr.scats['energy'][-1]) + int(rr.scats.offsets[-1])
        dtc3 = time.time() - t0
        row_bytes = (ec.row_words - 2) * 4 + 1
        drain = {'workload': 'C3 shape: Cu 20 keV DS tilt 70, scatter_output = type, position, energy; %d electrons in 4 '
                             'batches, Engine.run_batches(ordered="columns")' % (4 * nb),
                 'electrons_per_s': 4 * nb / dtc3, 'rows_per_s': rows / dtc3, 'row_bytes_over_pcie': row_bytes,
                 'd2h_GBps': (rows * row_bytes + 4 * nb * 40) / dtc3 / 1e9,
                 'note': 'wall clock incl. kernels, device regrouping (scan + columnar scatter) and D2H into pinned host '
                         'memory; PCIe-bound'}
        ec.close()

    if rank == 0:
        out = {
            'metric': 'electron trajectories/s', 'value': value, 'unit': 'electrons/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'events_per_s': events_s, 'events_per_electron': events_per_step / max(electrons_per_step, 1),
            'config': {'workload': 'C2', 'material': 'Si', 'E0_eV': 20000, 'Emin_eV': 10000, 'mode': 'DS',
                       's_tilt_deg': 70, 'tol_E': 1e-4, 'tol_W': 1e-7, 'maxScatt': 1000, 'num_el_per_gpu': n_el,
                       'outputs': 'one els row per electron + fused backscatter histograms + counters',
                       'cache': 'inputs larger than L2 (inverse-CDF tables %.2f GB resident in HBM)' % (table_bytes / 1e9),
                       'parallelism': 'electron ids partitioned across %d GPU(s); int64 histogram all-reduce' % world,
                       'pipelining': 'steps rotate over %d contexts / streams (shared tables); %d launches per step (main pass + tail hand-over passes)' % (NPIPE, launches)},
            'e2e': {'value': e2e, 'unit': 'electrons/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': args.steps * launches,
            'clocks': clocks,
            'roofline': roofline,
            'setup': {'tables_s': t_tables, 'table_shapes': {'Moller': [len(tabs['Moller'].Es), len(tabs['Moller'].Ws)],
                                                             'Gryz0': [len(tabs['Gryz'][0].Es), len(tabs['Gryz'][0].Ws)]},
                      'launch': eng.launch_info(), 'sm_mhz_est_from_mufu': mhz.value},
            'output_drain': drain,
            'physics': {'backscatter_fraction': eta, 'select_exact_frac': cnt['select_exact'] / max(events_per_step, 1),
                        'nan_clamped': cnt['nan_clamped']},
        }
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            el_s, ev_s, dtc, _ = cpu_sample(2000000, cores)
            out['cpu_baseline'] = {'value': el_s, 'unit': 'electrons/s', 'cores': cores, 'kind': 'port',
                                   'events_per_s': ev_s,
                                   'sample': '2e6 electrons of workload C2 (%.1f s), oracle C restatement of the '
                                             'reference loop, %d threads; tables at tol_E 1e-3 / tol_W 1e-5' % (dtc, cores)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
