#!/usr/bin/env python
"""bench.py -- headline benchmark of the per-electron trajectory loop (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

HEADLINE (config.workload = "C2", weak scaling): BASELINE.json configs[1] -- Si, E0 = 20 keV, DS mode (Rutherford +
Moller + Gryzinski + Quinn), default table tolerances (tol_E 1e-4, tol_W 1e-7), 70 deg tilt, num_el = 1e6 per GPU;
Emin = 10 keV and maxScatt = 1000 from the reference's inputFile/input.file (BASELINE.json leaves them open).  One step
= one pass of the trajectory kernel over 1e6 fresh electrons of the synthetic pencil beam (ids never repeat across
steps), histograms + counters fused, one els row per electron.  Every rank simulates its own 1e6 electrons per step
(disjoint Philox streams); the statistics accumulate ON THE DEVICE over the steps and the only collective is ONE int64
all-reduce of the histogram||counter block at the end of the run, inside the timed region.

`value` = electrons/s with tables resident in HBM, device-timed (CUDA events, max over ranks); `e2e` = the same through
the public call Engine.run_batches() with pinned HOST output buffers, i.e. including per step the launch-parameter
upload and the D2H copy of every els row, plus the single all-reduce and the D2H of the reduced block; `roofline` for
the trajectory kernel against the MEASURED FP32 / MUFU peaks of this GPU and the instruction-issue roof; `cpu_baseline`
= the oracle's C restatement of the reference loop on the host cores (bounded sample, rank 0).

NORTH-STAR LEGS (strong scaling: a FIXED job split over the N ranks by electron id, wall clock including the table
build, ONE all-reduce, SHA-1 of the reduced histogram||counter block -- equal digests at N = 1, 2, 4, 8 prove that the
result does not depend on the number of GPUs):
  target       C3 statistics: Cu, 20 keV, DS, tilt 70, 1e8 electrons, histograms + counters      (north_star target)
  c4           C4: Au, 30 keV, DS, tilt 0, 1e8 electrons, histograms only
  full_output  C3 with full scatter_output (type, position, energy) on a slice of the 1e8 ids, rows regrouped on the
               device into columns and drained to pinned host memory
  c5_cells     six cells of the C5 sweep (continuous JL / explicit and DS at 5 and 30 keV), 1e7 electrons each
  store        (rank 0, N = 1) HDF5 store write of a >= 1e6-electron run in the reference's layout
"""
import argparse
import hashlib
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
C3 = dict(WORKLOAD, material='Cu', num_el=10 ** 8)
C4 = dict(WORKLOAD, material='Au', E0=30000., Emin=15000., s_tilt=0., maxScatt=10000, num_el=10 ** 8, electron_output=[])
# algorithmic work per DS event: F = 181 + 24 S flops, M = 11 + 3 S MUFU ops, S = core shells
N_SHELLS = 3
FLOPS_PER_EVENT = 181 + 24 * N_SHELLS
MUFU_PER_EVENT = 11 + 3 * N_SHELLS
C5_CELLS = [('Al', 5000., 'cont', 'JL'), ('Cu', 5000., 'cont', 'explicit'), ('Si', 5000., 'DS', 'classical'),
            ('Au', 30000., 'cont', 'JL'), ('Si', 30000., 'cont', 'explicit'), ('Cu', 30000., 'DS', 'classical')]


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


def kernel_summary():
    """Per-event instruction count and DRAM traffic of the trajectory kernel from the latest `ncu --set full` capture
    (profiles/traj_kernel_summary.json, written by scripts/ncu_summary.py in the same gpurun call as the round's bench
    line).  `source_match` tells whether the capture was taken from the kernel sources of this checkout."""
    path = os.path.join(ROOT, 'profiles', 'traj_kernel_summary.json')
    if not os.path.exists(path):
        return None
    try:
        d = json.load(open(path))
    except Exception:
        return None
    d['source_match'] = (d.get('kernel_source_sha1') == kernel_source_sha1())
    return d


def kernel_source_sha1():
    h = hashlib.sha1()
    for f in ('mcd_traj.cu', 'mcd_device.cuh'):
        h.update(open(os.path.join(ROOT, 'mcdiscrete_b200', 'csrc', f), 'rb').read())
    return h.hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--num-el', type=int, default=NUM_EL)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--legs', default='target,c4,full_output,c5_cells,store',
                    help='north-star legs to run after the headline (comma separated; "none" = headline only)')
    ap.add_argument('--target-el', type=int, default=10 ** 8, help='electrons of the target / c4 legs (whole job)')
    ap.add_argument('--drain-el', type=int, default=12_500_000,
                    help='electrons of the full_output leg (whole job; 1.25e7 = one GPU\'s share of C3 on 8 GPUs)')
    ap.add_argument('--c5-el', type=int, default=10 ** 7, help='electrons per cell of the c5_cells leg (whole job)')
    ap.add_argument('--store-rows', type=int, default=50_000_000,
                    help='scats-row budget of the store leg (0 = every row of its 1e6 electrons: a 17 GB file)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    legs = [] if args.legs == 'none' else args.legs.split(',')

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
    from mcdiscrete_b200.distributed import allreduce_stats, rank_range, bind_to_gpu, restore_affinity
    # N > 1: every rank next to its own GPU (pinned host buffers on that NUMA node); one rank keeps the whole box
    binding = bind_to_gpu(local) if world > 1 else {'bound': False, 'why': 'single rank'}
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from mcdiscrete_b200 import _lib
    from mcdiscrete_b200.engine import Engine
    from mcdiscrete_b200.probTables import genTables
    import ctypes as C
    n_stats = _lib.HIST_TOTAL + C.sizeof(_lib.Counters) // 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def counters_of(block):
        """mcd_counters fields of a host int64 histogram||counter block"""
        c = block[_lib.HIST_TOTAL:]
        return {'electrons_done': int(c[0]), 'events': int(c[1]), 'outcome': {n: int(c[2 + i]) for i, n in enumerate(_lib.OUTCOMES)},
                'double_records': int(c[7]), 'scat_rows_stored': int(c[8]), 'scat_rows_dropped': int(c[9]),
                'nan_clamped': int(c[10]), 'select_exact': int(c[11])}

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
    f32p, mufu, mhz = C.c_float(), C.c_float(), C.c_float()
    _lib.check(eng.lib.mcd_measure_peaks(local, C.byref(f32p), C.byref(mufu), C.byref(mhz)))
    sm_count = C.c_int()
    _lib.check(eng.lib.mcd_device_info(local, None, 0, C.byref(sm_count), None, None, None))

    step_id = [0]

    def first_id():
        # fresh electrons every step, disjoint across ranks
        fid = (step_id[0] * world + rank) * n_el
        step_id[0] += 1
        return fid

    def device_step(k):
        # no reset between steps: the statistics of a context accumulate on the device over its steps
        e, ts = engs[k % NPIPE], tstreams[k % NPIPE]
        e.launch(first_id(), n_el, ts.cuda_stream)

    def reduce_once():
        """the ONE collective of a run: the contexts' blocks merged into the first (mcd_stats_merge, five launches of a
        1.6k-thread kernel), then all-reduced in place over the ranks"""
        cur = torch.cuda.current_stream()
        for ts in tstreams:
            cur.wait_stream(ts)
        for e in engs[1:]:
            eng.merge_stats_from(e, cur.cuda_stream)
        if world > 1:
            allreduce_stats(stats[0])

    def reset_all():
        torch.cuda.synchronize()
        for e, ts in zip(engs, tstreams):
            e.reset_stats(ts.cuda_stream)

    # ---- device-timed throughput ------------------------------------------------------------------------
    for k in range(args.warmup):
        device_step(k)
    reduce_once()
    barrier()
    reset_all()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    kernel_ms = []
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    ev0.record(tstreams[0])
    for ts in tstreams[1:]:
        ts.wait_event(ev0)                           # all streams start after the same timestamp
    for k in range(args.steps):
        device_step(k)
    reduce_once()                                    # current stream waits for every step, then the single all-reduce
    ev1.record(torch.cuda.current_stream())
    barrier()
    tw1 = time.time()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop(tw0, tw1)
    red = stats[0].cpu().numpy()
    cnt = counters_of(red)                           # of the K timed steps, summed over ranks
    events_total = cnt['events']
    electrons_total = cnt['electrons_done']
    assert electrons_total == n_el * world * args.steps, (electrons_total, n_el, world, args.steps)
    value = electrons_total / (ms * 1e-3)
    events_s = events_total / (ms * 1e-3)
    headline_sha = hashlib.sha1(red.tobytes()).hexdigest()[:16]

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
    hbm_peak, hbm_src = 6545.9, 'fallback (B200_PROFILING.md)'
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
        hbm_src = 'MEASURED_PEAKS.json'
    except Exception:
        pass
    ks = kernel_summary()
    traffic = issue = None
    clk_hz = 1e6 * float(clocks.get('sm_mhz') or clocks.get('sm_max_mhz') or mhz.value)
    issue_peak = 4.0 * sm_count.value * clk_hz           # warp instructions / s: 4 schedulers per SM, 1 per clock
    if ks:
        traffic = ks.get('dram_bytes_per_event', 0.0) * ev_local
        winst = ks.get('warp_inst_per_event')
        if winst:
            issue = {'warp_inst_per_event': winst, 'achieved_ginst_s': ev_rate * winst / 1e9, 'peak_ginst_s': issue_peak / 1e9,
                     'frac': ev_rate * winst / issue_peak,
                     'source': 'warp instructions per event: ncu capture %s (%s); peak = 4 schedulers x %d SMs x the SM '
                               'clock sampled in this run' % (ks.get('tag'), 'kernel sources unchanged since' if ks['source_match']
                                                              else 'STALE: kernel sources changed since', sm_count.value)}
    bound = 'mufu' if mufu_frac >= fp32_frac else 'fp32'
    roofline = {'bound': bound,
                'achieved': mufu_ach if bound == 'mufu' else fp32_ach,
                'peak': mufu.value if bound == 'mufu' else f32p.value,
                'unit': 'Ttranscendental/s' if bound == 'mufu' else 'TFLOP/s',
                'frac': max(mufu_frac, fp32_frac), 'traffic': traffic,
                'kernel': 'traj_kernel_pool<3> (main pass + %d tail hand-over passes of the same kernel)' % (launches - 1),
                'kernel_ms': step_ms, 'kernel_ms_standalone': k_ms, 'events_per_launch': ev_local,
                # scalars first (the driver's record keeps scalar sub-keys): the three compute roofs and HBM
                'mufu_frac': mufu_frac, 'fp32_frac': fp32_frac,
                'issue_frac': issue['frac'] if issue else None,
                'issue_warp_inst_per_event': issue['warp_inst_per_event'] if issue else None,
                'issue_peak_ginst_s': issue_peak / 1e9,
                'hbm_frac': (traffic / (step_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                'binding_roof': 'instruction issue' if issue and issue['frac'] > max(mufu_frac, fp32_frac) else bound,
                'traffic_source': ('ncu capture %s, %s' % (ks.get('tag'), 'same kernel sources' if ks['source_match'] else 'STALE kernel sources')) if ks else None,
                'timing': 'per-step device time inside the pipelined timed region; standalone = one step launched alone',
                'fp32': {'achieved_tflops': fp32_ach, 'peak_tflops': f32p.value, 'frac': fp32_frac,
                         'flops_per_event': FLOPS_PER_EVENT},
                'mufu': {'achieved_tops': mufu_ach, 'peak_tops': mufu.value, 'frac': mufu_frac,
                         'ops_per_event': MUFU_PER_EVENT},
                'issue': issue,
                'peak_source': 'measured on this GPU by mcd_measure_peaks (FFMA / MUFU.EX2 dependent chains); '
                               'MEASURED_PEAKS.json has no FP32/MUFU figure',
                'hbm': {'table_bytes_resident': table_bytes, 'peak_gbs': hbm_peak, 'peak_source': hbm_src,
                        'achieved_gbs': (traffic / (step_ms * 1e-3) / 1e9) if traffic else None,
                        'note': 'DRAM traffic of the trajectory kernel (ncu) over its CUDA-event duration; the kernel is '
                                'issue-bound, HBM is second order'}}

    # ---- end to end through the public API with pinned host outputs ------------------------------------------
    h2d = C.sizeof(_lib.Params) + 1024      # launch parameter block: there are no per-step input arrays
    d2h = n_el * 32 + (8 * n_stats) // max(args.steps, 1)

    e2e_marks = []

    def e2e_pass(nsteps):
        last = None
        del e2e_marks[:]
        eng.reset_pipeline_stats()
        for r in eng.run_batches(nsteps * n_el, n_el, first_id=first_id() + 10 ** 12, accumulate=True):
            last = r                                  # rows are in pinned host memory here
            _ = int(r.els['code'][0])
            e2e_marks.append(time.time())
        tot = eng.pipeline_stats()                    # sum over the pipeline's contexts ...
        e2e_marks.append(time.time())
        allreduce_stats(tot)                          # ... ONE all-reduce per run
        blk = tot.cpu().numpy()
        e2e_marks.append(time.time())
        return last, blk
    e2e_pass(NPIPE + 2)                               # warm every context of the pipeline (device buffers, pinned staging)
    barrier()
    t0 = time.time()
    r, e2e_block = e2e_pass(args.steps)
    t_done = time.time()
    barrier()
    dt = max_over_ranks(time.time() - t0)
    # where the end-to-end time goes (host clock, max over ranks): first batch on the host, steady interval between
    # batches, hand-over of the last batch .. merged statistics, the all-reduce + copy of the block, the final barrier
    mk = e2e_marks
    nb_ = len(mk) - 2
    e2e_detail = {
        'first_batch_ms': 1e3 * max_over_ranks(mk[0] - t0),
        'steady_ms_per_step': 1e3 * max_over_ranks((mk[nb_ - 1] - mk[0]) / max(nb_ - 1, 1)),
        'merge_stats_ms': 1e3 * max_over_ranks(mk[nb_] - mk[nb_ - 1]),
        'allreduce_and_copy_ms': 1e3 * max_over_ranks(mk[nb_ + 1] - mk[nb_]),
        'slowest_rank_total_ms': 1e3 * max_over_ranks(t_done - t0),
        'fastest_rank_total_ms': -1e3 * max_over_ranks(-(t_done - t0)),
        'wall_ms': 1e3 * dt}
    cnt_e2e = counters_of(e2e_block)
    assert cnt_e2e['electrons_done'] == n_el * world * args.steps
    e2e = n_el * world * args.steps / dt
    eta = cnt_e2e['outcome']['bks'] / cnt_e2e['electrons_done']
    for e in engs[1:]:
        e.close()

    # ---- north-star legs (strong scaling) ----------------------------------------------------------------------
    def strong_stats_leg(ipx, total, batch, els=False):
        """A fixed job of `total` electrons split over the ranks by id; statistics only; wall clock from before the
        engine is created (incl. the table build) to after the single all-reduce."""
        barrier()
        t0 = time.time()
        ex = Engine(ipx, device=local)
        if 'DS' in ipx['mode']:
            genTables(ipx, engine=ex)
        torch.cuda.synchronize()
        t_tab = time.time() - t0
        lo, n = rank_range(total, rank, world)
        t1 = time.time()
        for _ in ex.run_batches(n, batch, first_id=lo, accumulate=True, els=els):
            pass
        tot = ex.pipeline_stats()
        allreduce_stats(tot)
        blk = tot.cpu().numpy()
        barrier()
        t2 = time.time()
        wall, run = max_over_ranks(t2 - t0), max_over_ranks(t2 - t1)
        c = counters_of(blk)
        assert c['electrons_done'] == total, (c['electrons_done'], total)
        out = {'electrons': total, 'wall_s': wall, 'run_s': run, 'tables_s': max_over_ranks(t_tab),
               'electrons_per_s': total / run, 'events_per_s': c['events'] / run, 'events': c['events'],
               'stats_sha1': hashlib.sha1(blk.tobytes()).hexdigest()[:16], 'eta': c['outcome']['bks'] / total,
               'outcomes': c['outcome'], 'launch': ex.launch_info()}
        ex.close()
        return out

    leg = {}
    if 'target' in legs:
        leg['target'] = strong_stats_leg(dict(C3, electron_output=[]), args.target_el, 1_000_000)
        leg['target']['workload'] = 'C3 statistics: Cu 20 keV DS tilt 70 Emin 10 keV, default tolerances, histograms + counters'
    if 'c4' in legs:
        leg['c4'] = strong_stats_leg(C4, args.target_el, 500_000)
        leg['c4']['workload'] = 'C4: Au 30 keV DS tilt 0 Emin 15 keV maxScatt 10000, default tolerances, histograms only'
    if 'c5_cells' in legs:
        cells = []
        for name, E0, mode, model in C5_CELLS:
            emin = max(500., 0.05 * E0) if mode == 'cont' else E0 / 2
            ipx = dict(WORKLOAD, material=name, E0=E0, Emin=emin, mode=mode, Bethe_model=model, maxScatt=10000,
                       electron_output=[], num_el=args.c5_el)
            o_ = strong_stats_leg(ipx, args.c5_el, 1_000_000)
            cells.append({'cell': '%s %g keV %s' % (name, E0 / 1e3, 'DS' if mode == 'DS' else 'cont-' + model),
                          'wall_s': o_['wall_s'], 'run_s': o_['run_s'], 'electrons_per_s': o_['electrons_per_s'],
                          'events_per_s': o_['events_per_s'], 'eta': o_['eta'], 'stats_sha1': o_['stats_sha1']})
        leg['c5_cells'] = {'electrons_per_cell': args.c5_el, 'cells': cells,
                           'sha1_of_digests': hashlib.sha1(''.join(c['stats_sha1'] for c in cells).encode()).hexdigest()[:16]}
    if 'full_output' in legs:
        # C3 shape with full scatter_output: rows regrouped on the device into per-column arrays and drained to pinned
        # host memory, pipelined over two contexts; the slice [0, drain_el) of the ids is split over the ranks
        ipc = dict(C3, scatter_output=['type', 'position', 'energy'])
        ec = Engine(ipc, device=local)
        genTables(ipc, engine=ec)
        lo, n = rank_range(args.drain_el, rank, world)
        nb = max(1, min(625_000, (n + 1) // 2))
        for _ in ec.run_batches(min(n, 2 * nb), nb, first_id=10 ** 13 + lo, ordered='columns'):   # warm-up: pin the staging buffers
            pass
        ec._acc_host = None
        barrier()
        t0 = time.time()
        rows = 0
        for rr in ec.run_batches(n, nb, first_id=lo, ordered='columns', accumulate=True):
            rows += len(rr.scats)
            _ = float(rr.scats['energy'][-1]) + int(rr.scats.offsets[-1])
        tot = ec.pipeline_stats()
        allreduce_stats(tot)
        blk = tot.cpu().numpy()
        barrier()
        dtc3 = max_over_ranks(time.time() - t0)
        rows_t = torch.tensor([rows], device='cuda', dtype=torch.int64)
        if world > 1:
            dist.all_reduce(rows_t)
        rows_all = int(rows_t.item())
        c = counters_of(blk)
        assert c['electrons_done'] == args.drain_el and c['events'] == rows_all == c['scat_rows_stored']
        row_bytes = (ec.row_words - 2) * 4 + 1
        leg['full_output'] = {
            'workload': 'C3 with full scatter_output (type, position, energy): ids [0, %d) of the 1e8, '
                        'Engine.run_batches(ordered="columns")' % args.drain_el,
            'electrons': args.drain_el, 'wall_s': dtc3, 'electrons_per_s': args.drain_el / dtc3, 'rows': rows_all,
            'rows_per_s': rows_all / dtc3, 'row_bytes_over_pcie': row_bytes,
            'd2h_GBps': (rows_all * row_bytes + args.drain_el * 40) / dtc3 / 1e9,
            'stats_sha1': hashlib.sha1(blk.tobytes()).hexdigest()[:16],
            'seconds_for_1e8_at_this_rate': 1e8 / (args.drain_el / dtc3),
            'note': 'wall clock incl. kernels, device regrouping (scan + columnar scatter) and D2H into pinned host '
                    'memory on every rank; staging buffers pinned beforehand; tables resident'}
        ec.close()
    if 'store' in legs and world == 1:
        # the reference writes one pandas HDFStore through a listener process and "locks for largish data (electrons >
        # 500000)" (README.md:58-59).  Here: C2 with the reference's default outputs, 1e6 electrons, written by the
        # single-writer store (HDF5 layout of doScatter.py:63-67,112-117) while the GPU runs ahead.
        import shutil
        import tempfile
        from mcdiscrete_b200.store import ResultStore
        ips = dict(WORKLOAD, scatter_output=['type', 'position', 'energy'])
        es = Engine(ips, device=local)
        genTables(ips, engine=es)
        names = es.material.channel_names
        tmp = tempfile.mkdtemp(prefix='mcd_store_')
        path = os.path.join(tmp, 'c2.h5')
        n_store, nb = 1_000_000, 125_000
        budget = args.store_rows or (1 << 62)
        t0 = time.time()
        rows_w = els_w = 0
        try:
            with ResultStore(path, ips, 'hdf5') as st:
                fmt_used = st.fmt
                for rr in es.run_batches(n_store, nb, first_id=0, ordered='columns'):
                    if rows_w < budget:
                        st.append_result(rr, names)
                        rows_w += len(rr.scats)
                    else:                                   # past the row budget: the els table only
                        st.append_result(type(rr)(rr.first_id, rr.count, rr.els, {}, rr.counters, rr.kernel_ms), names)
                    els_w += rr.count
            dts = time.time() - t0
            size = os.path.getsize(path)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        leg['store'] = {'workload': 'C2 with scatter_output = type, position, energy (inputFile/input.file:37-39), 1e6 electrons',
                        'writer': fmt_used, 'els_rows_written': els_w, 'scats_rows_written': rows_w, 'file_GB': size / 1e9,
                        'wall_s': dts, 'rows_per_s': (rows_w + els_w) / dts, 'GBps_to_file': size / dts / 1e9,
                        'note': 'single writer, no lock: every els row of 1e6 electrons and %s; float64 columns and '
                                'an int64 index like the reference\'s DataFrames'
                                % ('every scats row' if not args.store_rows else 'the scats rows of the first batches up to '
                                   'the row budget (--store-rows 0 writes all of them)')}
        es.close()

    if rank == 0:
        tgt = leg.get('target', {})
        out = {
            'metric': 'electron trajectories/s', 'value': value, 'unit': 'electrons/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'events_per_s': events_s, 'events_per_electron': events_total / max(electrons_total, 1),
            'config': {'workload': 'C2', 'material': 'Si', 'E0_eV': 20000, 'Emin_eV': 10000, 'mode': 'DS',
                       's_tilt_deg': 70, 'tol_E': 1e-4, 'tol_W': 1e-7, 'maxScatt': 1000, 'num_el_per_gpu': n_el,
                       'outputs': 'one els row per electron + fused backscatter histograms + counters',
                       'cache': 'inputs larger than L2 (inverse-CDF tables %.2f GB resident in HBM)' % (table_bytes / 1e9),
                       'parallelism': 'electron ids partitioned across %d GPU(s); ONE int64 all-reduce per run' % world,
                       'pipelining': 'steps rotate over %d contexts / streams (shared tables); %d launches per step (main pass + tail hand-over passes + histogram pass over the els rows)' % (NPIPE, launches + 1),
                       # north-star legs, flattened (strong scaling: the same job at every N; equal digests = equal results)
                       'target_1e8_Cu_wall_s': tgt.get('wall_s'), 'target_1e8_Cu_run_s': tgt.get('run_s'),
                       'target_1e8_Cu_stats_sha1': tgt.get('stats_sha1'), 'target_electrons': tgt.get('electrons'),
                       'c4_1e8_Au_wall_s': leg.get('c4', {}).get('wall_s'), 'c4_1e8_Au_stats_sha1': leg.get('c4', {}).get('stats_sha1'),
                       'full_output_electrons_per_s': leg.get('full_output', {}).get('electrons_per_s'),
                       'full_output_rows_per_s': leg.get('full_output', {}).get('rows_per_s'),
                       'full_output_stats_sha1': leg.get('full_output', {}).get('stats_sha1'),
                       'c5_cells_sha1': leg.get('c5_cells', {}).get('sha1_of_digests')},
            'e2e': {'value': e2e, 'unit': 'electrons/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'e2e_detail': e2e_detail,
            'gpu_launches': args.steps * (launches + 1) + (NPIPE - 1),
            'clocks': clocks,
            'roofline': roofline,
            'setup': {'tables_s': t_tables, 'table_shapes': {'Moller': [len(tabs['Moller'].Es), len(tabs['Moller'].Ws)],
                                                             'Gryz0': [len(tabs['Gryz'][0].Es), len(tabs['Gryz'][0].Ws)]},
                      'launch': eng.launch_info(), 'sm_mhz_est_from_mufu': mhz.value, 'sm_count': sm_count.value},
            'physics': {'backscatter_fraction': eta, 'select_exact_frac': cnt['select_exact'] / max(events_total, 1),
                        'nan_clamped': cnt['nan_clamped'], 'headline_stats_sha1': headline_sha},
            'legs': leg,
        }
        out['host_binding'] = {k: v for k, v in binding.items() if not k.startswith('_')}
        if not args.no_cpu:
            restore_affinity(binding)                     # the CPU leg uses every core of the box
            cores = len(os.sched_getaffinity(0)) or 1
            el_s, ev_s, dtc, _ = cpu_sample(2000000, cores)
            out['cpu_baseline'] = {'value': el_s, 'unit': 'electrons/s', 'cores': cores, 'kind': 'port',
                                   'events_per_s': ev_s,
                                   'sample': '2e6 electrons of workload C2 (%.1f s), oracle C restatement of the '
                                             'reference loop, %d threads; tables at tol_E 1e-3 / tol_W 1e-5' % (dtc, cores)}
        # compact summary LAST (the driver keeps the tail of stdout)
        out['north_star'] = {
            'n_gpus': world,
            'target_1e8_Cu_DS_tilt70': {k: tgt.get(k) for k in ('electrons', 'wall_s', 'run_s', 'electrons_per_s', 'events_per_s', 'stats_sha1', 'eta')},
            'c4_1e8_Au': {k: leg.get('c4', {}).get(k) for k in ('electrons', 'wall_s', 'run_s', 'electrons_per_s', 'stats_sha1')},
            'full_output_slice': {k: leg.get('full_output', {}).get(k) for k in ('electrons', 'wall_s', 'rows_per_s', 'd2h_GBps', 'stats_sha1', 'seconds_for_1e8_at_this_rate')},
            'c5_cells_sha1': leg.get('c5_cells', {}).get('sha1_of_digests'),
            'roofs': {'mufu': mufu_frac, 'fp32': fp32_frac, 'issue': issue['frac'] if issue else None},
            'value_el_s': value, 'e2e_el_s': e2e}
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
