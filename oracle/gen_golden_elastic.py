"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the alternative elastic models of the path
(`elastic = Ruth_vanilla_wDefl | Ruth_nigram | Ruth_nigram_wDefl`), produced by the reference's
own unmodified modules under oracle/ref_shim.py, in the build container:

    python oracle/gen_golden_elastic.py      # writes tests/golden/elastic_formulas.json, traj_*_nigram*.npz

Same driver-side choices as oracle/gen_golden.py (Philox uniforms, float pickTable result).  What the reference
does with each label is RECORDED here, not assumed:
  * DS `Ruth_vanilla_wDefl` raises NameError (scattering.py:487 calls ruther_sigma_wDefl, which scattering.py:14
    does not import) -> only the formula crossSections.py:63-84 is pinned for it;
  * continuous mode tests substrings in the order 'vanilla', 'vanilla_wDefl', 'nigram', 'nigram_wDefl'
    (scattering.py:781-789), so the _wDefl labels run the plain forms; the golden trajectories show it.
"""
import json
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
    from tools import (material as rmat, crossSections as cs, scattering as sc, electron as rel, probTables as pt,
                       singleScatter as ss)

    # ---- formulas -------------------------------------------------------------------------------
    out = {'rows': [], 'reference_behaviour': {}}
    for name in ('Al', 'Si', 'Cu', 'Au'):
        Z = rmat.material(name).params['Z']
        for E in (500., 1234.5, 5000., 10000., 15000., 20000., 30000.):
            out['rows'].append(dict(material=name, Z=int(Z), E=E,
                                    vanilla=float(cs.ruther_sigma(E, Z)),
                                    vanilla_wDefl=float(cs.ruther_sigma_wDefl(E, Z)),
                                    nigram=float(cs.ruther_N_sigma(E, Z)),
                                    nigram_wDefl=float(cs.ruther_N_sigma_wDefl(E, Z))))

    # ---- DS trajectories on the committed Si_15k_small tables ------------------------------------------
    sc_pickTable = sc.pickTable
    sc.pickTable = lambda tab, en: (lambda r: (float(r[0]), r[1]))(sc_pickTable(tab, en))      # Q9
    ip0 = gg.base_input()
    tb = pt.genTables(ip0)
    g = np.load(os.path.join(GOLD, 'tables_Si_15k_small.npz'))
    assert np.array_equal(np.asarray(tb['Moller'].table.filled(0), np.float32), g['moller_table'])
    assert np.array_equal(np.asarray(tb['Gryz'][0].table.filled(0), np.float32), g['gryz0_table'])

    inj = gg.PhiloxInjector(gg.SEED, 'DS')
    orig_det = sc.scatter_discrete.det_type

    def det_type(self):
        orig_det(self)
        inj.last_type = self.type
    sc.scatter_discrete.det_type = det_type                     # observe only
    saved_random = random.random
    random.random = inj
    m = rmat.material('Si')
    names = ['Ruth', 'Moller'] + ['Gryz' + s for s in m.params['name_s']] + ['Quinn']
    t_ = np.radians(ip0['s_tilt'])
    dir0 = np.array([-np.sin(t_), 0., np.cos(t_)])

    def run_ds(elastic, ids):
        res = []
        for eid in ids:
            inj.start(eid)
            e = rel.electron(ip0['E0'], ip0['Emin'], np.zeros(3), dir0, {'el_output': gg.ALL_EL, 'scat_output': gg.ALL_SCAT})
            ss.trajectory_DS(e, m, ip0['Wc'], ip0['maxScatt'], ip0['thickness'], elastic, tb, ip0['diffMFP'])
            res.append((e.el_output.dict, e.scat_output.dict))
        return gg.flatten_traj(res, names)

    for elastic in ('Ruth_nigram', 'Ruth_nigram_wDefl'):
        np.savez_compressed(os.path.join(GOLD, 'traj_DS_Si_%s.npz' % elastic[5:]), ids=np.arange(20, 28),
                            **run_ds(elastic, range(20, 28)))
        out['reference_behaviour']['DS ' + elastic] = 'runs'
    try:
        run_ds('Ruth_vanilla_wDefl', [0])
        out['reference_behaviour']['DS Ruth_vanilla_wDefl'] = 'runs'
    except Exception as ex:                                      # NameError: ruther_sigma_wDefl
        out['reference_behaviour']['DS Ruth_vanilla_wDefl'] = '%s: %s' % (type(ex).__name__, ex)

    # ---- continuous trajectories (Al, classical Bethe) ----------------------------------------------
    inj.mode = 'cont'
    ipc = gg.base_input(material='Al', Emin=500., mode='cont', maxScatt=100000)
    mc = rmat.material('Al')
    cols = [c for c in gg.ALL_SCAT if c not in ('position', 'E')]

    def run_cont(elastic, ids):
        res = []
        for eid in ids:
            inj.start(eid)
            e = rel.electron(ipc['E0'], ipc['Emin'], np.zeros(3), dir0, {'el_output': gg.ALL_EL, 'scat_output': cols})
            ss.trajectory_cont_cl(e, mc, ipc['maxScatt'], ipc['thickness'], elastic, {})
            res.append((e.el_output.dict, e.scat_output.dict))
        return gg.flatten_traj(res, [elastic])

    van = np.load(os.path.join(GOLD, 'traj_cont_Al_classical.npz'))
    gn = run_cont('Ruth_nigram', range(6))
    np.savez_compressed(os.path.join(GOLD, 'traj_cont_Al_nigram.npz'), ids=np.arange(6), **gn)
    gnw = run_cont('Ruth_nigram_wDefl', range(6))
    gvw = run_cont('Ruth_vanilla_wDefl', range(6))
    same = lambda a, b: all(np.array_equal(a[k], b[k], equal_nan=True) for k in ('n_ev', 'outcome', 'final_E', 'pathl', 'energy'))
    out['reference_behaviour']['cont Ruth_nigram_wDefl'] = 'identical to Ruth_nigram' if same(gnw, gn) else 'differs'
    out['reference_behaviour']['cont Ruth_vanilla_wDefl'] = 'identical to Ruth_vanilla' if same(gvw, van) else 'differs'

    random.random = saved_random
    sc.scatter_discrete.det_type = orig_det
    sc.pickTable = sc_pickTable
    json.dump(out, open(os.path.join(GOLD, 'elastic_formulas.json'), 'w'), indent=1)
    print(json.dumps(out['reference_behaviour'], indent=1))


if __name__ == '__main__':
    main()
