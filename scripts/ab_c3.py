"""A/B timing of the row-recording path (C3 shape: Cu 20 keV DS 70 deg, scatter_output = type, position, energy):
trajectory kernel with row append, device regrouping (scan + row-major / columnar scatter), per library variant.

    python scripts/ab_c3.py [--n 1250000] [--workload Cu|Si|Au] variant.so ...

Every variant runs in its own process (MCD_LIB) on the same electron ids and prints CUDA-event times and a digest of
the regrouped columns; the digests of all variants must be equal."""
import argparse, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WL = {'Cu': dict(material='Cu', E0=20000., Emin=10000., s_tilt=70.),
      'Si': dict(material='Si', E0=20000., Emin=10000., s_tilt=70.),
      'Au': dict(material='Au', E0=30000., Emin=15000., s_tilt=0., maxScatt=10000)}


def child(a):
    import numpy as np
    import torch
    from mcdiscrete_b200.engine import Engine
    from mcdiscrete_b200.probTables import genTables
    import ctypes as C
    ip = dict(mode='DS', maxScatt=1000, thickness=3e6, Wc=10., tol_E=1e-4, tol_W=1e-7,
              electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'],
              scatter_output=a.cols.split(','), seed=12345)
    ip.update(WL[a.workload])
    eng = Engine(ip, device=0)
    genTables(ip, engine=eng)
    n = a.n
    eng.set_capacity(n, n * 400 + 65536)
    st = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    out = {'variant': os.environ.get('MCD_LIB', 'libmcd.so'), 'workload': a.workload, 'n': n, 'cols': a.cols}
    t_k, t_rows, t_cols = [], [], []
    for rep in range(a.reps + 1):
        eng.reset_stats(st)
        eng.launch(0, n, st)
        eng.sync(st)
        k_ms = eng.kernel_ms()
        ev[0].record()
        eng._ck(eng.lib.mcd_order_scats(eng.ctx, C.c_void_p(st)))
        ev[1].record()
        eng._ck(eng.lib.mcd_order_scats_columns(eng.ctx, C.c_void_p(st)))
        ev[2].record()
        torch.cuda.synchronize()
        if rep:
            t_k.append(k_ms); t_rows.append(ev[0].elapsed_time(ev[1])); t_cols.append(ev[1].elapsed_time(ev[2]))
    # (the order calls include a stream sync + the copy of the chunk counts: same overhead for every variant)
    els, scats = eng.drain(0, n, ordered='columns')
    h = hashlib.sha1()
    for k in sorted(scats.keys()):
        h.update(np.ascontiguousarray(scats[k]).tobytes())
    h.update(np.ascontiguousarray(els).tobytes())
    c = eng.counters()
    out.update(traj_ms=min(t_k), order_rows_ms=min(t_rows), order_cols_ms=min(t_cols), rows=int(c['scat_rows_stored']),
               events=int(c['events']), digest=h.hexdigest()[:16], launch=eng.launch_info())
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=1250000)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--workload', default='Cu')
    ap.add_argument('--cols', default='type,position,energy')
    ap.add_argument('--child', action='store_true')
    ap.add_argument('variants', nargs='*')
    a = ap.parse_args()
    if a.child:
        return child(a)
    for v in a.variants or ['']:
        env = dict(os.environ)
        if v:
            env['MCD_LIB'] = os.path.abspath(v)
        subprocess.run([sys.executable, os.path.abspath(__file__), '--child', '--n', str(a.n), '--reps', str(a.reps),
                        '--workload', a.workload, '--cols', a.cols], env=env)


if __name__ == '__main__':
    main()
