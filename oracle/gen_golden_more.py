"""TEST INFRASTRUCTURE ONLY.  Further golden vectors from the reference's own, unmodified code (oracle/ref_shim.py),
for cases the first set does not reach:

    python oracle/gen_golden_more.py     # writes tests/golden/traj_DS_Si_thin.npz, traj_DS_Au.npz (+ its tables)

  * a thin film (thickness 9000 A, normal incidence): 'trsm' outcomes of update_xyz (electron.py:73-81);
  * Au (8 core shells, 30 keV): the widest channel list of the materials the reference ships; tables restricted to
    the two the hot path reads (Moller + outermost Gryzinski), like gen_golden.py does for Cu.
Same driver-side choices as oracle/gen_golden.py (Philox uniforms, float pickTable result).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg              # noqa: E402
import ref_shim                      # noqa: E402

GOLD = gg.GOLD


def main():
    import random
    ref_shim.load_reference()
    from tools import material as rmat, scattering as sc, electron as rel, probTables as pt, singleScatter as ss
    sc_pickTable = sc.pickTable
    sc.pickTable = lambda tab, en: (lambda r: (float(r[0]), r[1]))(sc_pickTable(tab, en))      # Q9
    inj = gg.PhiloxInjector(gg.SEED, 'DS')
    orig_det = sc.scatter_discrete.det_type

    def det_type(self):
        orig_det(self)
        inj.last_type = self.type
    sc.scatter_discrete.det_type = det_type                     # observe only
    saved_random = random.random
    random.random = inj

    def run_ds(ip, tb, ids):
        m = rmat.material(ip['material'])
        t_ = np.radians(ip['s_tilt'])
        dir0 = np.array([-np.sin(t_), 0., np.cos(t_)])
        res = []
        for eid in ids:
            inj.start(eid)
            e = rel.electron(ip['E0'], ip['Emin'], np.zeros(3), dir0, {'el_output': gg.ALL_EL, 'scat_output': gg.ALL_SCAT})
            ss.trajectory_DS(e, m, ip['Wc'], ip['maxScatt'], ip['thickness'], ip['elastic'], tb, ip['diffMFP'])
            res.append((e.el_output.dict, e.scat_output.dict))
        names = ['Ruth', 'Moller'] + ['Gryz' + s for s in m.params['name_s']] + ['Quinn']
        return gg.flatten_traj(res, names)

    def tables_arrays(tb):
        arrs = {}
        for nm, T in (('moller', tb['Moller']), ('gryz0', tb['Gryz'][0])):
            arrs[nm + '_Ws'] = np.asarray(T.Ws)
            arrs[nm + '_Es'] = np.asarray(T.Es)
            arrs[nm + '_table'] = np.asarray(T.table.filled(0), dtype=np.float32)
        return arrs

    # ---- thin film: the committed Si_15k_small tables, thickness 9000 A, normal incidence ---------------------
    ip = gg.base_input(thickness=9000., s_tilt=0.)
    tb = pt.genTables(ip)
    g = np.load(os.path.join(GOLD, 'tables_Si_15k_small.npz'))
    assert np.array_equal(np.asarray(tb['Moller'].table.filled(0), np.float32), g['moller_table'])
    ids = np.arange(40, 70)
    out = run_ds(ip, tb, ids)
    print('thin film outcomes:', np.bincount(out['outcome'], minlength=5))
    np.savez_compressed(os.path.join(GOLD, 'traj_DS_Si_thin.npz'), ids=ids, **out)

    # ---- Au, 30 keV -> 27 keV, normal incidence; hot tables only ----------------------------------------------
    ip = gg.base_input(material='Au', E0=30000., Emin=27000., s_tilt=0.)
    m = rmat.material('Au')
    kw = dict(E_range=(ip['Emin'], ip['E0']), tol_E=ip['tol_E'], tol_W=ip['tol_W'], mat=m,
              mapTarget='../tables/integration', chunk_size=100)
    T = []
    for t in (pt.probTable(type='Moller', shell=m.params['name_val'], func=pt.moller_dCS, Wc=ip['Wc'], **kw),
              pt.probTable(type='Gryzinski', shell=m.params['name_s'][-1], func=pt.gryz_dCS, **kw)):
        t.generate(); t.mapToMemory(); t.readFromMemory()
        T.append(t)
    tb = {'Moller': T[0], 'Gryz': [T[1]]}
    ids = np.arange(3)
    out = run_ds(ip, tb, ids)
    print('Au events per electron:', out['n_ev'], 'channels used:', np.unique(out['type']))
    np.savez_compressed(os.path.join(GOLD, 'traj_DS_Au.npz'), ids=ids, **out, **tables_arrays(tb))

    random.random = saved_random
    sc.scatter_discrete.det_type = orig_det
    sc.pickTable = sc_pickTable


if __name__ == '__main__':
    main()
