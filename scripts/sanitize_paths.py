"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_paths.py
pool kernel + hand-over passes + histogram pass, row-recording kernel, both regrouping layouts (6-, 3- and 14-word
rows: the last one needs the opt-in 56 KB of dynamic shared memory), continuous kernels, Mott run, launch tunings."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mcdiscrete_b200.engine import Engine
from mcdiscrete_b200.probTables import genTables
from mcdiscrete_b200.extFunctions import mottTable

ALL = ['type', 'position', 'pathl', 'E', 'E_loss', 'pol_angle', 'az_angle', 'energy', 'direction']
base = dict(E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, mode='DS', maxScatt=1000, Wc=10.,
            tol_E=1e-2, tol_W=1e-3, electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'], seed=3)
n = 3000
ref = None
for cols in ([], ['type', 'position', 'energy'], ['type', 'energy'], ALL):
    eng = Engine(dict(base, scatter_output=cols), device=0)
    genTables(dict(base, scatter_output=cols), engine=eng)
    for mode in ((False,) if not cols else (True, 'columns')):
        r = eng.run(count=n, first_id=5, ordered=mode)
        if ref is None:
            ref = r.els.copy()
        assert np.array_equal(r.els, ref)
        if cols:
            assert len(r.scats) == r.counters['events']
    if cols:
        tot = sum(len(b.scats) for b in eng.run_batches(n, 1000, first_id=5, ordered='columns'))
        assert tot == r.counters['events']
    eng.close()
for kw in (dict(handover_passes=1), dict(handover_passes=4, handover_threshold=7), dict(statistics_kernel=1)):
    eng = Engine(dict(base, scatter_output=[]), device=0)
    eng.tuning(**kw)
    genTables(dict(base, scatter_output=[]), engine=eng)
    assert np.array_equal(eng.run(count=n, first_id=5).els, ref)
    eng.close()
for model in ('classical', 'JL', 'explicit'):
    ipc = dict(base, material='Al', Emin=2000., mode='cont', Bethe_model=model, maxScatt=100000, scatter_output=['energy'])
    eng = Engine(ipc, device=0)
    r = eng.run(count=500, first_id=0, ordered='columns')
    assert len(r.scats) == r.counters['events']
    eng.close()
ipm = dict(base, elastic='Mott', scatter_output=['type', 'pol_angle'],
           mott_path=os.path.join(ROOT, 'tests', 'data', 'tables', 'mott') + os.sep)
eng = Engine(ipm, device=0)
genTables(ipm, engine=eng)
r = eng.run(count=1000, first_id=0, ordered=True)
assert (np.asarray(r.scats['type']) == 0).any()
eng.close()
print('sanitize_paths ok')
