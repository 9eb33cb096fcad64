"""Run the BASELINE.json configurations (scaled to one GPU where noted) and print one JSON line each:
electrons/s, events/s, kernel ms, outcome fractions, drain GB/s.  Used to fill BASELINE.md section 4.

    python scripts/run_configs.py [C1 C2 C3 C4 C5 ...]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                    # noqa: E402
from mcdiscrete_b200.engine import Engine             # noqa: E402
from mcdiscrete_b200.probTables import genTables      # noqa: E402

BASE = dict(thickness=3e6, elastic='Ruth_vanilla', Bethe_model='classical', diffMFP=False, gen_tables=True,
            tol_E=1e-4, tol_W=1e-7, Wc=10., seed=12345,
            electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'], scatter_output=[])


def run(tag, ip, n, batch=None, scats=False, reps=2):
    import torch
    eng = Engine(ip, device=0)
    t0 = time.time()
    if ip['mode'] == 'DS':
        genTables(ip, engine=eng)
    torch.cuda.synchronize()
    t_tab = time.time() - t0
    batch = batch or n
    # warm-up: first touch of the pinned staging buffers (page-locking GBs takes seconds) and of the kernel
    res = eng.run(count=min(batch, n), first_id=10 ** 10, pinned=True, drain=False)
    eng.drain(10 ** 10, min(batch, n), want_els=True, want_scats=scats, pinned=True)
    tot_ev = tot_el = 0
    k_ms = 0.0
    drain_s = 0.0
    drain_bytes = 0
    out_counts = np.zeros(5, np.int64)
    t0 = time.time()
    for rep in range(reps):
        for first in range(0, n, batch):
            cnt = min(batch, n - first)
            res = eng.run(count=cnt, first_id=rep * n + first, pinned=True, drain=False)
            k_ms += res.kernel_ms
            tot_ev += res.counters['events']
            tot_el += cnt
            out_counts += np.array([res.counters['outcome'][k] for k in ('abs', 'bks', 'trsm', 'trdf', 'long')])
            t1 = time.time()
            els, sc = eng.drain(rep * n + first, cnt, want_els=True, want_scats=scats, pinned=True)
            drain_s += time.time() - t1
            drain_bytes += cnt * 32 + (len(sc) * eng.row_words * 4 if len(sc) else 0)
    wall = time.time() - t0
    out = dict(config=tag, material=ip['material'], E0=ip['E0'], Emin=ip['Emin'], mode=ip['mode'],
               bethe=ip.get('Bethe_model'), tilt=ip['s_tilt'], electrons=tot_el, events_per_electron=tot_ev / tot_el,
               kernel_electrons_per_s=tot_el / (k_ms * 1e-3), kernel_events_per_s=tot_ev / (k_ms * 1e-3),
               e2e_electrons_per_s=tot_el / wall, kernel_ms_total=k_ms, tables_s=t_tab,
               drain_GBps=(drain_bytes / 1e9) / max(drain_s, 1e-9), drain_bytes=drain_bytes,
               outcome_frac={k: float(v) / tot_el for k, v in zip(('abs', 'bks', 'trsm', 'trdf', 'long'), out_counts)},
               select_exact_frac=res.counters['select_exact'] / max(res.counters['events'], 1),
               nan_clamped=res.counters['nan_clamped'], launch=eng.launch_info())
    print(json.dumps(out), flush=True)
    eng.close()


def run_pipelined(tag, ip, total, batch, ordered=True):
    """Full scatter_output through the pipelined public API: kernel k+1 overlaps the D2H drain of batch k; rows
    arrive on the host already regrouped per electron (device-side scan + scatter)."""
    import torch
    eng = Engine(ip, device=0)
    genTables(ip, engine=eng)
    for r in eng.run_batches(2 * batch, batch, first_id=10 ** 10, ordered=ordered):   # warm-up: pin staging buffers
        pass
    torch.cuda.synchronize()
    t0 = time.time()
    rows = ev = 0
    kms = 0.0
    for r in eng.run_batches(total, batch, first_id=0, ordered=ordered):
        rows += len(r.scats)
        ev += r.counters['events']
        kms += r.kernel_ms
        _ = float(r.scats['energy'][-1]) + int(r.scats.offsets[-1])
    wall = time.time() - t0
    row_bytes = eng.row_words * 4 if ordered is True else (eng.row_words - 2) * 4 + 1
    nbytes = rows * row_bytes + total * 32 + (total + 1) * 8
    print(json.dumps(dict(config=tag, material=ip['material'], electrons=total, batch=batch, scats_rows=rows,
                          row_bytes=row_bytes, e2e_electrons_per_s=total / wall, e2e_rows_per_s=rows / wall,
                          e2e_GBps_to_host=nbytes / 1e9 / wall, kernel_events_per_s=ev / (kms * 1e-3),
                          wall_s=wall, kernel_s=kms * 1e-3)), flush=True)
    eng.close()


def main():
    want = sys.argv[1:] or ['C1', 'C2', 'C3', 'C4', 'C5']
    if 'C1' in want:   # Al, cont classical, 1e3 electrons (the reference's own CPU-sized case) + a GPU-sized batch
        ip = dict(BASE, material='Al', E0=20000., Emin=500., mode='cont', s_tilt=70., maxScatt=100000, num_el=1000)
        run('C1 (1e3)', ip, 1000)
        run('C1 (1e6)', ip, 1000000)
    if 'C2' in want:
        ip = dict(BASE, material='Si', E0=20000., Emin=10000., mode='DS', s_tilt=70., maxScatt=1000, num_el=10 ** 6)
        run('C2', ip, 10 ** 6)
    if 'C3' in want:   # Cu full scatter_output: 1e8 over 8 GPUs = 1.25e7 per GPU; here 2.5e6 in batches of 1.25e6
        ip = dict(BASE, material='Cu', E0=20000., Emin=10000., mode='DS', s_tilt=70., maxScatt=1000, num_el=10 ** 8,
                  scatter_output=['type', 'position', 'energy'])
        run_pipelined('C3 full scatter_output, 5e6 e- of the 1.25e7 per-GPU share, pipelined + device regroup', ip, 5000000, 1250000)
        run_pipelined('C3 full scatter_output, 5e6 e-, pipelined + COLUMNAR device regroup (no id words)', ip, 5000000, 1250000,
                      ordered='columns')
        ip2 = dict(ip, scatter_output=[])
        run('C3 kernel only (1e7, no scats)', ip2, 10 ** 7, reps=1)
    if 'C4' in want:   # Au 30 keV tilt 0 histograms only; Emin = E0/2, maxScatt 10000 (stated; BASELINE leaves them open)
        ip = dict(BASE, material='Au', E0=30000., Emin=15000., mode='DS', s_tilt=0., maxScatt=10000, num_el=10 ** 8)
        run('C4 (1.25e7 = per-GPU share of 1e8)', ip, 12500000, batch=12500000, reps=1)
    if 'C5' in want:   # sweep: E0 x material x {JL, explicit, DS}; 1e6 electrons each here (1e7 in BASELINE)
        for mat in ('Al', 'Si', 'Cu', 'Au'):
            for E0 in (5000., 10000., 15000., 20000., 25000., 30000.):
                for mode, bm in (('cont', 'JL'), ('cont', 'explicit'), ('DS', 'classical')):
                    emin = max(500., 0.05 * E0) if mode == 'cont' else E0 / 2
                    ip = dict(BASE, material=mat, E0=E0, Emin=emin, mode=mode, Bethe_model=bm, s_tilt=70.,
                              maxScatt=10000, num_el=10 ** 6, tol_E=1e-3, tol_W=1e-5)
                    try:
                        run('C5', ip, 10 ** 6, reps=1)
                    except Exception as e:          # noqa: BLE001
                        print(json.dumps(dict(config='C5', material=mat, E0=E0, mode=mode, bethe=bm, error=str(e))), flush=True)


if __name__ == '__main__':
    main()
