"""Why does a DS cell run on the plain kernel?  Prints what mcd_create decided (MCD_DEBUG=1): static channel order,
accuracy of the per-cell records at every grid resolution."""
import os, sys
os.environ['MCD_DEBUG'] = '1'
sys.path.insert(0, '.')
from mcdiscrete_b200.engine import Engine
for name, E0 in [(a.split(':')[0], float(a.split(':')[1])) for a in sys.argv[1:]]:
    ip = dict(E0=E0, Emin=E0 / 2, material=name, s_tilt=70., thickness=3e6, mode='DS', maxScatt=10000, Wc=10.,
              electron_output=['outcome'], scatter_output=[], seed=1)
    print('==', name, E0, flush=True)
    sys.stderr.flush()
    eng = Engine(ip, device=0)
    print('   kernel', eng.launch_info(), flush=True)
    eng.close()
