"""torchrun script: N ranks simulate disjoint id ranges of one run and all-reduce the statistics block;
rank 0 re-runs ALL ids on its own GPU and checks that histograms and counters are EQUAL (Philox streams are
keyed on the electron id, so the result must not depend on the partition).  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                         # noqa: E402
import torch                                               # noqa: E402
import torch.distributed as dist                           # noqa: E402
from mcdiscrete_b200.engine import Engine                  # noqa: E402
from mcdiscrete_b200.probTables import genTables           # noqa: E402
from mcdiscrete_b200.distributed import init_from_env, run_partitioned   # noqa: E402

rank, local, world = init_from_env()
torch.cuda.set_device(local)
ip = dict(num_el=400000, E0=20000., Emin=10000., material='Cu', s_tilt=70., thickness=3e6, elastic='Ruth_vanilla',
          mode='DS', Bethe_model='classical', diffMFP=False, gen_tables=True, tol_E=1e-3, tol_W=1e-5, maxScatt=1000,
          Wc=10., electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'], scatter_output=[], seed=2026)
eng = Engine(ip, device=local)
genTables(ip, engine=eng)
res, cnt, hist = run_partitioned(eng, ip['num_el'], first_id=0)
els_all = [None] * world
if world > 1:
    dist.all_gather_object(els_all, res.els.copy())
else:
    els_all = [res.els.copy()]
if rank == 0:
    solo = eng.run(count=ip['num_el'], first_id=0)
    h1 = eng.histograms()
    c1 = eng.counters()
    same_hist = all(np.array_equal(hist[k][0], h1[k][0]) for k in hist)
    same_cnt = all(cnt[k] == c1[k] for k in ('electrons_done', 'events', 'outcome', 'double_records', 'nan_clamped'))
    same_rows = np.array_equal(np.concatenate(els_all), solo.els)
    print(json.dumps(dict(world=world, electrons=cnt['electrons_done'], events=cnt['events'],
                          histograms_equal=bool(same_hist), counters_equal=bool(same_cnt),
                          els_rows_equal=bool(same_rows), eta=cnt['outcome']['bks'] / cnt['electrons_done'])))
    assert same_hist and same_cnt and same_rows
if world > 1:
    dist.destroy_process_group()
