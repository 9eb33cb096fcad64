"""Same-seed and independent-seed comparison of one continuous / DS cell, GPU vs the float64 C oracle (diagnostic)."""
import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, mcd_oracle as o, mcd_oracle_cbind as oc
from mcdiscrete_b200.engine import Engine
name, E0, mode, model = sys.argv[1], float(sys.argv[2]), sys.argv[3], sys.argv[4]
Emin = max(500., 0.05 * E0) if mode == 'cont' else E0 / 2
ip = dict(E0=E0, Emin=Emin, material=name, s_tilt=70., thickness=3e6, mode=mode, Bethe_model=model, maxScatt=10000,
          Wc=10., tol_E=1e-3, tol_W=1e-5, electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'],
          scatter_output=['energy'], seed=5150, elastic='Ruth_vanilla')
tabs = o.gen_tables(ip, strict=False) if mode == 'DS' else None
eng = Engine(ip, device=0)
if tabs:
    eng.table_upload(0, tabs['Moller'].Es, tabs['Moller'].Ws, tabs['Moller'].table)
    eng.table_upload(1, tabs['Gryz'][0].Es, tabs['Gryz'][0].Ws, tabs['Gryz'][0].table)
    eng.tables = tabs
n = 100000
nt = os.cpu_count()
res = eng.run(count=n, first_id=0, ordered=True)
same = oc.run(ip, 0, n, 5150, tabs, nthreads=nt)
nev = np.diff(res.scats.offsets.astype(np.int64))
dE = np.abs(res.els['final_E'].astype(np.float64) - same['final_E'])
print(name, E0, mode, model, 'same seed: outcome mismatches', int((res.outcome != same['outcome']).sum()),
      'event-count mismatches', int((nev != same['n_ev']).sum()), '|dE|>1eV', int((dE > 1).sum()),
      'mean nev gpu %.3f oracle %.3f' % (nev.mean(), same['n_ev'].mean()),
      'eta gpu %.4f oracle %.4f' % ((res.outcome == 1).mean(), (same['outcome'] == 1).mean()))
bad = np.flatnonzero(nev != same['n_ev'])[:8]
print(' first mismatching ids', bad, nev[bad], same['n_ev'][bad], res.outcome[bad], same['outcome'][bad])
for seed in (4243, 2025, 1, 2, 3):
    r = oc.run(ip, 0, 40000, seed, tabs, nthreads=nt)
    print(' oracle seed %d: mean nev %.3f (se %.3f) eta %.4f' % (seed, r['n_ev'].mean(), r['n_ev'].std() / 200, (r['outcome'] == 1).mean()))
