"""TEST INFRASTRUCTURE ONLY.  STATISTICAL golden samples from the reference's own, unmodified code WITH ITS OWN RNG
(`random`, Mersenne Twister), in the build container:

    python oracle/gen_golden_stats.py [n_DS n_cont [config ...]]      # writes tests/golden/stats_*.npz (default 8000 / 6000 electrons; float32 is enough for the distributions)

The other golden files replace `random.random` by Philox streams so that single trajectories can be compared event
for event.  These samples do not: every electron is `random.seed(SEED + id)` followed by the reference's
`trajectory_DS` / `trajectory_cont_cl` (singleScatter.py:10-41, 205-239) exactly as `multiTraj_*` calls them
(multiScatter.py:97-112), so the distributions they hold (outcome, final energy, exit direction, last position,
trajectory length) are the reference's, with nothing of this repository's random-number plumbing in the loop.  The
C port, and on the GPU box the CUDA path, are compared with them as INDEPENDENT samples (backscatter coefficient
within 3 sigma, two-sample KS): tests/test_oracle_c.py, tests/test_gpu_statistics.py.

Configurations: the committed small-tolerance tables (tests/golden/tables_*.npz are the same objects), so that the
reference, the ports and the GPU all sample the same tables.  Same driver-side choices as gen_golden.py for what is
not random: float pickTable result (Q9), Cu tables through `probTable` for the two hot tables (Q3).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg              # noqa: E402
import ref_shim                      # noqa: E402
import mcd_oracle as port            # noqa: E402

GOLD = gg.GOLD
SEED = 20261018


def main():
    import random
    n_ds = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
    n_cont = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
    ref_shim.load_reference()
    from tools import (material as rmat, scattering as sc, electron as rel, probTables as pt, singleScatter as ss)
    sc_pickTable = sc.pickTable
    sc.pickTable = lambda tab, en: (lambda r: (float(r[0]), r[1]))(sc_pickTable(tab, en))      # Q9

    def hot_tables(ip):
        m = rmat.material(ip['material'])
        kw = dict(E_range=(ip['Emin'], ip['E0']), tol_E=ip['tol_E'], tol_W=ip['tol_W'], mat=m,
                  mapTarget='../tables/integration', chunk_size=100)
        out = []
        for T in (pt.probTable(type='Moller', shell=m.params['name_val'], func=pt.moller_dCS, Wc=ip['Wc'], **kw),
                  pt.probTable(type='Gryzinski', shell=m.params['name_s'][-1], func=pt.gryz_dCS, **kw)):
            T.generate(); T.mapToMemory(); T.readFromMemory()
            out.append(T)
        return {'Moller': out[0], 'Gryz': [out[1]]}

    def collect(els_dicts, nev, secs, ip):
        oc = np.array([port.OUTCOMES.index(d['outcome'][0]) for d in els_dicts], np.int8)
        return dict(outcome=oc, final_E=np.array([float(d['final_E'][0]) for d in els_dicts], np.float32),
                    final_dir=np.array([[d['final_dir'][k][0] for k in ('dx', 'dy', 'dz')] for d in els_dicts], np.float32),
                    last_pos=np.array([[d['last_pos'][k][0] for k in ('last_x', 'last_y', 'last_z')] for d in els_dicts], np.float32),
                    n_ev=np.array(nev, np.int32), double_records=np.array([len(d['outcome']) - 1 for d in els_dicts], np.int8),
                    seed=np.array(SEED),
                    input=np.array(repr({k: ip[k] for k in ('material', 'E0', 'Emin', 's_tilt', 'mode', 'Bethe_model',
                                                             'maxScatt', 'thickness', 'Wc', 'tol_E', 'tol_W', 'elastic')})))

    def run(ip, n, tables):
        m = rmat.material(ip['material'])
        t_ = np.radians(ip['s_tilt'])
        dir0 = np.array([-np.sin(t_), 0., np.cos(t_)])           # multiScatter.py:97-100
        out, nev = [], []
        t0 = time.time()
        for eid in range(n):
            random.seed(SEED + eid)                              # (the reference: random.seed(pid * time), :94)
            e = rel.electron(ip['E0'], ip['Emin'], np.zeros(3), dir0, {'el_output': gg.ALL_EL, 'scat_output': ['energy']})
            if ip['mode'] == 'DS':
                ss.trajectory_DS(e, m, ip['Wc'], ip['maxScatt'], ip['thickness'], ip['elastic'], tables, ip['diffMFP'])
            else:
                ss.trajectory_cont_cl(e, m, ip['maxScatt'], ip['thickness'], ip['elastic'], tables)
            out.append(e.el_output.dict)
            nev.append(len(e.scat_output.dict['energy']))
        r = collect(out, nev, time.time() - t0, ip)
        r['_seconds'] = time.time() - t0          # printed only: the file must regenerate bit for bit
        return r

    cfgs = [('DS_Si', gg.base_input(), n_ds, 'Si_15k_small'),
            ('DS_Cu', gg.base_input(material='Cu', Emin=10000.), n_ds // 2, 'Cu_10k_small'),
            ('cont_Al', gg.base_input(material='Al', Emin=500., mode='cont', maxScatt=100000), n_cont, None),
            # a 9000 Angstrom film at normal incidence ('trsm' outcomes, electron.py:73-81) and the Mott elastic channel
            # on the synthetic ioffe-layout table of oracle/gen_golden_mott.py
            ('DS_Si_thin', gg.base_input(thickness=9000., s_tilt=0.), n_ds // 2, 'Si_15k_small'),
            ('DS_Si_mott', gg.base_input(elastic='Mott'), n_ds // 2, 'Si_15k_small')]
    only = [a for a in sys.argv[3:]]
    if only:
        cfgs = [c for c in cfgs if c[0] in only]
    import shutil
    work_mott = os.path.abspath(os.path.join('..', 'tables', 'mott', 'ioffe'))
    os.makedirs(work_mott, exist_ok=True)
    shutil.copy(os.path.join(os.path.dirname(HERE), 'tests', 'data', 'tables', 'mott', 'ioffe', 'Si.ioffe'), work_mott)
    for tag, ip, n, tkey in cfgs:
        tb = None
        if tkey:
            tb = hot_tables(ip) if ip['material'] == 'Cu' else pt.genTables(ip)
            g = np.load(os.path.join(GOLD, 'tables_%s.npz' % tkey))
            assert np.array_equal(np.asarray(tb['Moller'].table.filled(0), np.float32), g['moller_table'])
            assert np.array_equal(np.asarray(tb['Gryz'][0].table.filled(0), np.float32), g['gryz0_table'])
        r = run(ip, n, tb)
        secs = r.pop('_seconds')
        np.savez_compressed(os.path.join(GOLD, 'stats_%s.npz' % tag), **r)
        print('%s: %d electrons in %.0f s, eta %.4f, mean events %.1f' % (tag, n, secs,
                                                                         (r['outcome'] == 1).mean(), r['n_ev'].mean()), flush=True)
    sc.pickTable = sc_pickTable


if __name__ == '__main__':
    main()
