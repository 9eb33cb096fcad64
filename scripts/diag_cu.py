import os, sys, math
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np
from scipy import stats
import mcd_oracle as o, mcd_oracle_cbind as oc
from mcdiscrete_b200.engine import Engine
for name in ('Cu', 'Si'):
    ip = dict(E0=20000., Emin=10000., material=name, s_tilt=70., thickness=3e6, mode='DS', maxScatt=4000, Wc=10.,
              tol_E=1e-3, tol_W=1e-4, electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'], scatter_output=[], seed=777)
    tabs = o.gen_tables(ip, strict=False)
    eng = Engine(ip, device=0)
    eng.table_upload(0, tabs['Moller'].Es, tabs['Moller'].Ws, tabs['Moller'].table)
    eng.table_upload(1, tabs['Gryz'][0].Es, tabs['Gryz'][0].Ws, tabs['Gryz'][0].table)
    eng.tables = tabs
    n = 100000
    res = eng.run(count=n, first_id=0)
    same = oc.run(ip, 0, n, 777, tabs, nthreads=os.cpu_count())
    out = res.outcome
    agree = out == same['outcome']
    print(name, 'same-seed outcome agreement', agree.mean(), 'eta gpu', (out == 1).mean(), 'eta orc', (same['outcome'] == 1).mean())
    dE = res.els['final_E'][agree].astype(np.float64) - same['final_E'][agree]
    print('  final_E diff among agreeing: mean %.4g max %.4g; frac exactly-equal-ev-count n/a' % (dE.mean(), np.abs(dE).max()))
    for seed in (4242, 1, 2):
        ref = oc.run(ip, 0, 40000, seed, tabs, nthreads=os.cpu_count())
        b, rb = out == 1, ref['outcome'] == 1
        rng = np.random.default_rng(0)
        a = res.els['final_E'][b].astype(np.float64); c = ref['final_E'][rb]
        p_raw = stats.ks_2samp(a, c).pvalue
        p_jit = stats.ks_2samp(a + rng.uniform(-.05, .05, len(a)), c + rng.uniform(-.05, .05, len(c))).pvalue
        p_z = stats.ks_2samp(res.els['last_z'][out == 0].astype(np.float64), ref['last_pos'][ref['outcome'] == 0, 2]).pvalue
        # oracle vs oracle (same-seed oracle sample as "gpu")
        p_oo = stats.ks_2samp(same['final_E'][same['outcome'] == 1], c).pvalue
        p_zoo = stats.ks_2samp(same['last_pos'][same['outcome'] == 0, 2], ref['last_pos'][ref['outcome'] == 0, 2]).pvalue
        print('  seed', seed, 'p final_E raw %.4f jitter %.4f | abs z %.4f | oracle-vs-oracle final_E %.4f abs z %.4f' % (p_raw, p_jit, p_z, p_oo, p_zoo))
    eng.close()
