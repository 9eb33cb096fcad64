"""One batch of the C3 workload (Cu, 20 keV, DS, 70 deg, full scatter_output) for ncu: trajectory kernel with row
append + device regrouping (scan + scatter)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcdiscrete_b200.engine import Engine
from mcdiscrete_b200.probTables import genTables
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1250000
ip = dict(material='Cu', E0=20000., Emin=10000., mode='DS', s_tilt=70., maxScatt=1000, thickness=3e6, Wc=10.,
          tol_E=1e-4, tol_W=1e-7, electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'],
          scatter_output=['type', 'position', 'energy'], seed=12345)
eng = Engine(ip, device=0)
genTables(ip, engine=eng)
for rep, mode in enumerate((True, True, 'columns', 'columns')):     # row-major regrouping, then the columnar one
    r = eng.run(count=n, first_id=(rep % 2) * n, pinned=True, ordered=mode)
    print(rep, mode, len(r.scats), r.kernel_ms)
