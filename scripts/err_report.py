"""Print the measured max relative error of the FP32 device formulas against the float64 oracle."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, mcd_oracle as o
from mcdiscrete_b200.engine import Engine
rng = np.random.default_rng(1)
for name in ('Al', 'Si', 'Cu', 'Au'):
    ip = dict(E0=30000., Emin=200., material=name, s_tilt=70., maxScatt=1000, mode='cont')
    eng = Engine(ip, device=0)
    m = o.Material(name)
    E = rng.uniform(200., 30000., 200000).astype(np.float32)
    sig, mfp = eng.test_sigmas(E)
    ref = np.array([o.sigmas_ds(float(e), m, 10.) for e in E[:20000]])
    nz = ref > 0
    rel = np.abs(sig[:20000].astype(np.float64) - ref) / np.where(nz, ref, 1)
    rel[~nz] = 0
    ref_mfp = np.array([o.mfp_from_sigma(sum(r), m.atnd) for r in ref])
    print(name, 'sigma max rel err per channel', ['%.2e' % x for x in rel.max(axis=0)], 'mfp %.2e' % (np.abs(mfp[:20000] - ref_mfp) / ref_mfp).max())
    u = np.clip(rng.random(20000).astype(np.float32), 2 ** -24, 1 - 2 ** -24)
    c2, om = eng.test_ruth(E[:20000], u)
    ref_om = np.array([o.alpha(float(e), m.Z) * float(v) / (1. + o.alpha(float(e), m.Z) - float(v)) for e, v in zip(E[:20000], u)])
    print('   ruth om max rel %.2e' % (np.abs(om - ref_om) / ref_om).max())
    W = (u * E[:20000] * 0.49).astype(np.float32)
    c2b, omb = eng.test_bincoll(E[:20000], W)
    refb = np.array([o.binary_coll(float(e), float(w), m.fermi_e) for e, w in zip(E[:20000], W)])
    ok = ~np.isnan(refb) & ~np.isnan(c2b)
    print('   bincoll c2 max rel %.2e' % (np.abs(c2b[ok] - refb[ok]) / refb[ok]).max())
    eng.close()
