"""Statistical parity report at larger sizes than the test-suite uses: GPU (2e6 electrons) vs the float64 C oracle
(2e5 electrons, INDEPENDENT Philox seed) on the same tables; plus same-seed agreement on 1e5 electrons.
Writes one JSON line per configuration (profiles/r1_statistics.jsonl).
--default-tol: BASELINE's default tolerances (tol_E 1e-4, tol_W 1e-7); the GPU runs on tables BUILT ON THE DEVICE, the
oracle on its own tables (about 30 s of numpy per material)."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
from scipy import stats
import mcd_oracle as o, mcd_oracle_cbind as oc
from mcdiscrete_b200.engine import Engine

CFG = [('C2 Si 20keV DS 70deg', dict(material='Si', E0=20000., Emin=10000., s_tilt=70., mode='DS', maxScatt=1000)),
       ('C3 Cu 20keV DS 70deg', dict(material='Cu', E0=20000., Emin=10000., s_tilt=70., mode='DS', maxScatt=1000)),
       ('C4 Au 30keV DS 0deg', dict(material='Au', E0=30000., Emin=15000., s_tilt=0., mode='DS', maxScatt=10000)),
       ('C1 Al 20keV cont-classical 70deg', dict(material='Al', E0=20000., Emin=500., s_tilt=70., mode='cont', Bethe_model='classical', maxScatt=100000)),
       ('C5 Cu 15keV cont-JL 70deg', dict(material='Cu', E0=15000., Emin=750., s_tilt=70., mode='cont', Bethe_model='JL', maxScatt=100000)),
       ('C5 Au 10keV cont-explicit 70deg', dict(material='Au', E0=10000., Emin=500., s_tilt=70., mode='cont', Bethe_model='explicit', maxScatt=100000))]
NT = os.cpu_count() or 1
DEFAULT_TOL = '--default-tol' in sys.argv
for tag, kw in CFG:
    ip = dict(thickness=3e6, Wc=10., tol_E=1e-3, tol_W=1e-4, Bethe_model='classical', seed=20261017,
              electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'], scatter_output=[])
    ip.update(kw)
    if DEFAULT_TOL:
        ip.update(tol_E=1e-4, tol_W=1e-7)
    tabs = o.gen_tables(ip, strict=False) if ip['mode'] == 'DS' else None
    eng = Engine(ip, device=0)
    if tabs and DEFAULT_TOL:
        from mcdiscrete_b200.probTables import genTables
        dt = genTables(ip, engine=eng)
        assert np.array_equal(dt['Moller'].Es, tabs['Moller'].Es) and np.array_equal(dt['Gryz'][0].Ws, tabs['Gryz'][0].Ws)
    elif tabs:
        eng.table_upload(0, tabs['Moller'].Es, tabs['Moller'].Ws, tabs['Moller'].table)
        eng.table_upload(1, tabs['Gryz'][0].Es, tabs['Gryz'][0].Ws, tabs['Gryz'][0].table)
        eng.tables = tabs
    n_gpu, n_ref = 2000000, 200000
    res = eng.run(count=n_gpu, first_id=0)
    t0 = time.time()
    ref = oc.run(ip, 0, n_ref, 987654321, tabs, nthreads=NT)
    t_ref = time.time() - t0
    same = oc.run(ip, 0, 100000, ip['seed'], tabs, nthreads=NT)
    out = res.outcome
    eg, er = (out == 1).mean(), (ref['outcome'] == 1).mean()
    sig = math.sqrt(eg * (1 - eg) / n_gpu + er * (1 - er) / n_ref)
    b, rb = out == 1, ref['outcome'] == 1
    rng = np.random.default_rng(0)
    jit = lambda x: np.asarray(x, np.float64) + rng.uniform(-0.05, 0.05, len(x))
    ks = {
        'final_E': stats.ks_2samp(jit(res.els['final_E'][b]), jit(ref['final_E'][rb])).pvalue,
        'exit_polar': stats.ks_2samp(np.arccos(np.clip(-res.els['dz'][b].astype(np.float64), -1, 1)), np.arccos(np.clip(-ref['final_dir'][rb, 2], -1, 1))).pvalue,
        'exit_azimuth': stats.ks_2samp(np.arctan2(res.els['dy'][b], res.els['dx'][b]).astype(np.float64), np.arctan2(ref['final_dir'][rb, 1], ref['final_dir'][rb, 0])).pvalue,
        'depth_last_z': stats.ks_2samp(res.els['last_z'][b].astype(np.float64), ref['last_pos'][rb, 2]).pvalue,
        'abs_last_z': stats.ks_2samp(res.els['last_z'][out == 0].astype(np.float64), ref['last_pos'][ref['outcome'] == 0, 2]).pvalue,
    }
    print(json.dumps(dict(config=tag, tol_E=ip['tol_E'], tol_W=ip['tol_W'], tables='device-built vs oracle-built' if (tabs and DEFAULT_TOL) else 'oracle-built, uploaded', n_gpu=n_gpu, n_oracle=n_ref, eta_gpu=float(eg), eta_oracle=float(er), eta_diff_in_sigma=float((eg - er) / sig),
                          ks_p={k: float(v) for k, v in ks.items()}, events_per_electron_gpu=res.counters['events'] / n_gpu,
                          events_per_electron_oracle=float(ref['n_ev'].mean()),
                          same_seed_outcome_agreement=float((out[:100000] == same['outcome']).mean()),
                          same_seed_event_count_agreement=None, oracle_seconds=t_ref, oracle_threads=NT,
                          nan_clamped=res.counters['nan_clamped'], select_exact=res.counters['select_exact'])), flush=True)
    eng.close()
