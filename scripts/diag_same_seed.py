import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, mcd_oracle as o, mcd_oracle_cbind as oc
from mcdiscrete_b200.engine import Engine
ip = dict(E0=20000., Emin=10000., material='Si', s_tilt=70., thickness=3e6, mode='DS', maxScatt=1000, Wc=10.,
          tol_E=1e-3, tol_W=1e-4, electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'], scatter_output=['type'], seed=777)
tabs = o.gen_tables(ip, strict=False)
eng = Engine(ip, device=0)
eng.table_upload(0, tabs['Moller'].Es, tabs['Moller'].Ws, tabs['Moller'].table)
eng.table_upload(1, tabs['Gryz'][0].Es, tabs['Gryz'][0].Ws, tabs['Gryz'][0].table)
eng.tables = tabs
n = 200000
res = eng.run(count=n, first_id=0, ordered=True)
same = oc.run(ip, 0, n, 777, tabs, nthreads=os.cpu_count())
nev = np.diff(res.scats.offsets.astype(np.int64))
dE = np.abs(res.els['final_E'].astype(np.float64) - same['final_E'])
print('electrons', n, 'outcome mismatches', int((res.outcome != same['outcome']).sum()), 'event-count mismatches', int((nev != same['n_ev']).sum()),
      '|dE|>1eV', int((dE > 1).sum()), '|dE|>0.1', int((dE > 0.1).sum()), 'select_exact', res.counters['select_exact'])
bad = np.flatnonzero(nev != same['n_ev'])[:10]
print('first mismatching ids', bad, nev[bad], same['n_ev'][bad], res.outcome[bad], same['outcome'][bad])
