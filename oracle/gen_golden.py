"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/* by running the reference's own,
unmodified modules (under oracle/ref_shim.py) in the build container.

    python oracle/gen_golden.py          # rewrites tests/golden/*.npz / *.json

The reference ships no golden vectors; these files ARE the pin for the
oracle port (oracle/mcd_oracle.py, oracle/mcd_oracle_c.c) and, through it, for the CUDA path.
Nothing here is restated physics: every number comes out of /root/reference code.

Driver-side environment choices (documented deviations, none touches a reference file):
  * ``random.random`` is replaced by a Philox4x32-10 stream keyed on (seed, electron id,
    event index) -- the reference seeds ``random`` with pid*time per job (multiScatter.py:94,
    SURVEY Q8), which is not reproducible.
  * ``scattering.pickTable`` results are cast to python float (numpy-1.x-era float64 energy,
    SURVEY Q9).
  * continuous JL / explicit: ``scatter_continuous_classical`` is subclassed to call the
    reference's ``stoppingPowers.bethe_mod_sp_k`` / ``bethe_mod_sp`` (the shipped JL/explicit
    classes call the classical form and multiTraj_cont crashes).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim                      # noqa: E402
import mcd_oracle as port            # noqa: E402  (only for the Philox stream definition)

GOLD = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
SEED = 12345
ALL_SCAT = ['type', 'position', 'pathl', 'E', 'E_loss', 'pol_angle', 'az_angle', 'energy', 'direction']
ALL_EL = ['outcome', 'final_E', 'final_dir', 'last_pos']


class PhiloxInjector:
    """Feeds the reference's random.random() calls: DS events draw (type, path, [channel], phi),
    continuous events draw (path, polar, phi)."""

    def __init__(self, seed, mode):
        self.seed, self.mode = seed, mode
        self.eid = self.ev = self.k = 0
        self.last_type = None

    def start(self, eid):
        self.eid, self.ev, self.k = eid, 0, 0

    def __call__(self):
        u = port.event_uniforms(self.seed, self.eid, self.ev)
        if self.mode == 'DS':
            if self.k == 2 and self.last_type == 'Quinn':
                self.k = 3
            val = u[self.k]
            self.k += 1
            if self.k == 4:
                self.k, self.ev = 0, self.ev + 1
        else:
            val = u[self.k]
            self.k += 1
            if self.k == 3:
                self.k, self.ev = 0, self.ev + 1
        return val


def base_input(**kw):
    d = dict(num_el=10, E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6,
             elastic='Ruth_vanilla', mode='DS', Bethe_model='classical', diffMFP=False, gen_tables=True,
             tol_E=1e-2, tol_W=1e-3, maxScatt=1000, Wc=10., electron_output=ALL_EL, scatter_output=ALL_SCAT)
    d.update(kw)
    return d


def flatten_traj(results_per_el, chan_names):
    """list of (els dict, scats dict) -> flat arrays."""
    out = {k: [] for k in ('n_ev', 'n_els', 'outcome', 'final_E', 'final_dir', 'last_pos', 'type', 'position',
                           'pathl', 'E', 'E_loss', 'pol_angle', 'az_angle', 'energy', 'direction')}
    for els, sc in results_per_el:
        out['n_els'].append(len(els['outcome']))
        out['outcome'] += [port.OUTCOMES.index(o) for o in els['outcome']]
        out['final_E'] += [float(v) for v in els['final_E']]
        fd = els['final_dir']
        out['final_dir'] += list(zip(fd['dx'], fd['dy'], fd['dz']))
        lp = els['last_pos']
        out['last_pos'] += list(zip(lp['last_x'], lp['last_y'], lp['last_z']))
        n = len(sc['energy'])
        out['n_ev'].append(n)
        out['type'] += [chan_names.index(t) if t in chan_names else 0 for t in sc['type'][:n]]
        if 'position' in sc:
            p = sc['position']
            out['position'] += list(zip(p['x'][:n], p['y'][:n], p['z'][:n]))   # drop dangling 'long' position (Q6)
        for k in ('pathl', 'E', 'E_loss', 'pol_angle', 'az_angle', 'energy'):
            if k in sc:
                out[k] += [float(v) for v in sc[k][:n]]
        out['direction'] += [np.asarray(v, float) for v in sc['direction'][:n]]
    return {k: np.asarray(v) for k, v in out.items()}


def main():
    import random
    t = ref_shim.load_reference()
    from tools import (material as rmat, crossSections as cs, stoppingPowers as sp, scattering as sc,
                       electron as rel, probTables as pt, singleScatter as ss, multiScatter as ms)
    os.makedirs(GOLD, exist_ok=True)
    rng = np.random.default_rng(SEED)

    # ---- 1. formulas -------------------------------------------------------------------
    formulas = {}
    for name in ('Al', 'Si', 'Cu', 'Au'):
        m = rmat.material(name)
        P = m.params
        energies = [100000., 30000., 20000., 15000., 10000., 9000., 5000., 2000., 1000., 500., 200., 120.]
        rows = []
        for E in energies:
            sig = [cs.ruther_sigma(E, P['Z']), cs.moller_sigma(E, 10., P['n_val'])]
            sig += [cs.gryz_sigma(E, P['Es'][s], P['ns'][s]) for s in P['name_s']]
            sig.append(cs.quinn_sigma(E, m.plasmon_e, m.fermi_e, m.atnd))
            sig = [float(x) for x in sig]
            rows.append(dict(
                E=E, sigma=sig, mfp_total=float(sc.mfp_from_sigma(sum(sig), m.atnd)),
                mfp_ruth=float(sc.mfp_from_sigma(sig[0], m.atnd)),
                sp_classical=float(sp.bethe_cl_sp(P['Z'], E, m.atnd)),
                sp_JL=float(sp.bethe_mod_sp_k(P['Z'], E, m.atnd, P['bethe_k'])),
                sp_explicit=float(sp.bethe_mod_sp(E, m.atnd, [P['ns'][s] for s in P['name_s']],
                                                  [P['Es'][s] for s in P['name_s']], P['n_val'], P['E_val']))))
        formulas[name] = dict(atnd=float(m.atnd), fermi_e=float(m.fermi_e), plasmon_e=float(m.plasmon_e),
                              Z=P['Z'], rows=rows)
    json.dump(formulas, open(os.path.join(GOLD, 'formulas.json'), 'w'), indent=1)

    # ---- 2. channel pick (pickFromSigmas) ------------------------------------------------
    sig_sets, us, picks = [], [], []
    names7 = ['Ruth', 'Moller', 'Gryz1s', 'Gryz2s', 'Gryz2p', 'Quinn', 'pad']
    for trial in range(400):
        n = int(rng.integers(2, 12))
        s = np.exp(rng.uniform(-3, 3, n))
        if trial % 5 == 0:
            s[rng.integers(0, n)] = 0.0
        if trial % 7 == 0 and n > 2:
            i, j = rng.choice(n, 2, replace=False)
            s[i] = s[j]                                      # exact tie -> stable order matters
        if trial % 11 == 0:
            s[s < 0.3] = 0.0
        if s.sum() == 0:
            s[0] = 1.0
        keys = ['c%02d' % i for i in range(n)]
        d = dict(zip(keys, [float(x) for x in s]))
        u = float(rng.random())
        if trial % 13 == 0:                                  # u exactly on a boundary of the sorted cumsum
            srt = sorted(d.values(), reverse=True)
            pr = np.cumsum(srt) / sum(srt)
            u = float(pr[rng.integers(0, len(pr) - 1)]) if len(pr) > 1 else u
        with ref_shim.injected_uniforms([u]):
            k = sc.pickFromSigmas(d)
        row = np.full(11, -1.0)
        row[:n] = s
        sig_sets.append(row)
        us.append(u)
        picks.append(keys.index(k))
    np.savez_compressed(os.path.join(GOLD, 'select.npz'), sigma=np.array(sig_sets), u=np.array(us),
                        pick=np.array(picks))

    # ---- 3. angles and rotation ----------------------------------------------------------
    n = 300
    E = rng.uniform(100., 30000., n)
    Z = rng.choice([13, 14, 29, 79], n)
    u = rng.random(n)
    u[:6] = [0.0, 1e-12, 0.5, 1 - 1e-9, 0.999999, 2.0 ** -24]
    c2_ruth = []
    for i in range(n):
        with ref_shim.injected_uniforms([u[i]]):
            c2_ruth.append(float(sc.Rutherford_halfPol(float(E[i]), int(Z[i]))))
    W = rng.uniform(0, 1, n) * E * 0.55
    Ef = rng.choice([11.6587, 12.46422, 34.8428, 27.3364], n)
    with np.errstate(all='ignore'):
        c2_bin = [float(sc.binaryCollModel(float(E[i]), float(W[i]), float(Ef[i]))) for i in range(n)]
    # rotation: random orthonormal (d, y), random c2/halfPhi incl. edge values
    d_in, y_in, c2_in, hp_in, d_out, y_out = [], [], [], [], [], []
    for i in range(n):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        y = np.cross(d, rng.normal(size=3))
        y /= np.linalg.norm(y)
        c2 = float(rng.random()) if i % 3 else float(1 - 10 ** rng.uniform(-9, -1))
        hp = float(np.pi * rng.random())
        if i < 4:
            c2, hp = [(1.0, 0.3), (1.0, 2.5), (0.5, 0.0), (0.999, np.pi / 2)][i]
        e = rel.electron(20000., 10000., np.zeros(3), d.copy(), {'el_output': [], 'scat_output': []})
        e.y_local = y.copy()
        e.update_direction(c2, hp)
        d_in.append(d); y_in.append(y); c2_in.append(c2); hp_in.append(hp)
        d_out.append(np.asarray(e.dir, float)); y_out.append(np.asarray(e.y_local, float))
    np.savez_compressed(os.path.join(GOLD, 'angles.npz'), E=E, Z=Z, u=u, c2_ruth=np.array(c2_ruth),
                        W=W, Ef=Ef, c2_bin=np.array(c2_bin), d_in=np.array(d_in), y_in=np.array(y_in),
                        c2_in=np.array(c2_in), hp_in=np.array(hp_in), d_out=np.array(d_out),
                        y_out=np.array(y_out))

    # ---- 4. tables -----------------------------------------------------------------------
    sc_pickTable = sc.pickTable
    sc.pickTable = lambda tab, en: (lambda r: (float(r[0]), r[1]))(sc_pickTable(tab, en))      # Q9
    table_cfgs = {
        'Si_15k_small': base_input(),
        'Al_500_small': base_input(material='Al', Emin=500., tol_E=1e-2, tol_W=1e-3),
        'Cu_10k_small': base_input(material='Cu', Emin=10000., tol_E=1e-2, tol_W=1e-3),
    }
    tabs = {}
    axes_sizes = {}

    def hot_tables(ip):
        """genTables (probTables.py:541-614) restricted to the two tables the hot path reads
        (Moller and the outermost Gryzinski shell = tables['Gryz'][0], scattering.py:625): for Cu
        the reference's own genTables dies with StopIteration in the never-read 1s table."""
        m = rmat.material(ip['material'])
        kw = dict(E_range=(ip['Emin'], ip['E0']), tol_E=ip['tol_E'], tol_W=ip['tol_W'], mat=m,
                  mapTarget='../tables/integration', chunk_size=100)
        out = []
        for T in (pt.probTable(type='Moller', shell=m.params['name_val'], func=pt.moller_dCS, Wc=ip['Wc'], **kw),
                  pt.probTable(type='Gryzinski', shell=m.params['name_s'][-1], func=pt.gryz_dCS, **kw)):
            T.generate(); T.mapToMemory(); T.readFromMemory()
            out.append(T)
        return {'Moller': out[0], 'Gryz': [out[1]]}

    def detach(tb):
        """The reference maps '<type>_<shell>.table' files (probTables.py:513-535); a later genTables for
        the same material overwrites them under the live memmap.  Keep an in-memory masked copy."""
        import numpy.ma as ma
        for T in [tb['Moller']] + list(tb['Gryz']):
            T.table = ma.masked_array(np.array(T.table.filled(0)), mask=np.array(ma.getmaskarray(T.table)))
        return tb

    for key, ip in table_cfgs.items():
        tb = detach(hot_tables(ip) if ip['material'] == 'Cu' else pt.genTables(ip))
        tabs[key] = tb
        arrs = {}
        for nm, T in (('moller', tb['Moller']), ('gryz0', tb['Gryz'][0])):
            arrs[nm + '_Ws'] = np.asarray(T.Ws)
            arrs[nm + '_Es'] = np.asarray(T.Es)
            arrs[nm + '_table'] = np.asarray(T.table.filled(0), dtype=np.float32)
        arrs['gryz0_shell'] = np.array(tb['Gryz'][0].shell)
        # pickTable lookups
        m = rmat.material(ip['material'])
        nl = 400
        Es_q = rng.uniform(ip['Emin'] * 1.0001, ip['E0'], nl)
        Es_q[:3] = [ip['E0'], float(tb['Moller'].Es[3]), float(tb['Moller'].Es[1]) + 1e-3]
        us_q = rng.random(nl)
        us_q[:5] = [0.0, 1e-9, 0.999999, 0.5, float(tb['Moller'].table[5][7])]
        for nm, T in (('moller', tb['Moller']), ('gryz0', tb['Gryz'][0])):
            Wq = []
            for Eq, uq in zip(Es_q, us_q):
                with ref_shim.injected_uniforms([uq]):
                    Wq.append(sc.pickTable(T, float(Eq))[0])
            arrs[nm + '_lookup_W'] = np.array(Wq)
        arrs['lookup_E'] = Es_q
        arrs['lookup_u'] = us_q
        np.savez_compressed(os.path.join(GOLD, 'tables_%s.npz' % key), **arrs)
        axes_sizes[key] = dict(input={k: ip[k] for k in ('material', 'E0', 'Emin', 'tol_E', 'tol_W', 'Wc')},
                               moller=[len(tb['Moller'].Es), len(tb['Moller'].Ws)],
                               gryz=[[g.shell, len(g.Es), len(g.Ws)] for g in tb['Gryz']])
    # axis sizes only, at tighter tolerances (tables too big to commit)
    for key, ip in {'Si_10k_mid': base_input(Emin=10000., tol_E=1e-3, tol_W=1e-5),
                    'Al_500_mid': base_input(material='Al', Emin=500., tol_E=1e-3, tol_W=1e-4)}.items():
        tb = pt.genTables(ip)
        axes_sizes[key] = dict(input={k: ip[k] for k in ('material', 'E0', 'Emin', 'tol_E', 'tol_W', 'Wc')},
                               moller=[len(tb['Moller'].Es), len(tb['Moller'].Ws)],
                               gryz=[[g.shell, len(g.Es), len(g.Ws)] for g in tb['Gryz']])
        T, G = tb['Moller'], tb['Gryz'][0]
        rows_m = [0, 1, len(T.Es) // 2, len(T.Es) - 1]
        rows_g = [0, 1, len(G.Es) // 2, len(G.Es) - 1]
        np.savez_compressed(os.path.join(GOLD, 'tablerows_%s.npz' % key),
                            moller_rows_idx=np.array(rows_m), gryz0_rows_idx=np.array(rows_g),
                            moller_rows=np.asarray(T.table.filled(0)[rows_m], dtype=np.float32),
                            gryz0_rows=np.asarray(G.table.filled(0)[rows_g], dtype=np.float32))
    json.dump(axes_sizes, open(os.path.join(GOLD, 'table_axes.json'), 'w'), indent=1)

    # ---- 5. full DS trajectories under Philox uniforms -------------------------------------
    inj = PhiloxInjector(SEED, 'DS')
    orig_det = sc.scatter_discrete.det_type

    def det_type(self):
        orig_det(self)
        inj.last_type = self.type
    sc.scatter_discrete.det_type = det_type                     # observe only
    saved_random = random.random
    random.random = inj

    def run_ds(ip, tb, ids):
        m = rmat.material(ip['material'])
        pos0 = np.array([0., 0., 0.])
        t_ = np.radians(ip['s_tilt'])
        dir0 = np.array([-np.sin(t_), 0., np.cos(t_)])           # multiScatter.py:97-100
        res = []
        for eid in ids:
            inj.start(eid)
            e = rel.electron(ip['E0'], ip['Emin'], pos0, dir0, {'el_output': ALL_EL, 'scat_output': ALL_SCAT})
            ss.trajectory_DS(e, m, ip['Wc'], ip['maxScatt'], ip['thickness'], ip['elastic'], tb, ip['diffMFP'])
            res.append((e.el_output.dict, e.scat_output.dict))
        names = ['Ruth', 'Moller'] + ['Gryz' + s for s in m.params['name_s']] + ['Quinn']
        return flatten_traj(res, names)

    np.savez_compressed(os.path.join(GOLD, 'traj_DS_Si.npz'), ids=np.arange(10),
                        **run_ds(table_cfgs['Si_15k_small'], tabs['Si_15k_small'], range(10)))
    ip = dict(table_cfgs['Cu_10k_small'])
    np.savez_compressed(os.path.join(GOLD, 'traj_DS_Cu.npz'), ids=np.arange(100, 106),
                        **run_ds(ip, tabs['Cu_10k_small'], range(100, 106)))
    # Q6: 'long'
    ip = dict(table_cfgs['Si_15k_small'], maxScatt=5)
    np.savez_compressed(os.path.join(GOLD, 'traj_DS_Si_long.npz'), ids=np.arange(4),
                        **run_ds(ip, tabs['Si_15k_small'], range(4)))
    # Q4: bks then abs double record (tilt 89.5, Emin just under E0); small tables for this Emin
    ip = base_input(s_tilt=89.5, Emin=19985.)
    tb = detach(pt.genTables(ip))
    g = run_ds(ip, tb, range(300))
    arrs = {}
    for nm, T in (('moller', tb['Moller']), ('gryz0', tb['Gryz'][0])):
        arrs[nm + '_Ws'] = np.asarray(T.Ws); arrs[nm + '_Es'] = np.asarray(T.Es)
        arrs[nm + '_table'] = np.asarray(T.table.filled(0), dtype=np.float32)
    np.savez_compressed(os.path.join(GOLD, 'traj_DS_Si_q4.npz'), ids=np.arange(300),
                        n_ev=g['n_ev'], n_els=g['n_els'], outcome=g['outcome'], final_E=g['final_E'],
                        final_dir=g['final_dir'], last_pos=g['last_pos'], **arrs)
    print('Q4 double records:', int((g['n_els'] == 2).sum()), 'of', len(g['n_els']))

    # ---- 6. continuous trajectories ---------------------------------------------------------
    inj.mode = 'cont'

    class _JL(sc.scatter_continuous_classical):
        def compute_Eloss(self):
            self.E_loss = self.pathl * sp.bethe_mod_sp_k(self.m_Z, self.Ei, self.m_atnd, self._mat.params['bethe_k'])
            self.scat_output.addToList('E_loss', self.E_loss)

    class _EX(sc.scatter_continuous_classical):
        def compute_Eloss(self):
            P = self._mat.params
            self.E_loss = self.pathl * sp.bethe_mod_sp(self.Ei, self.m_atnd, [P['ns'][s] for s in P['name_s']],
                                                       [P['Es'][s] for s in P['name_s']], P['n_val'], P['E_val'])
            self.scat_output.addToList('E_loss', self.E_loss)

    def run_cont(ip, ids, model):
        m = rmat.material(ip['material'])
        t_ = np.radians(ip['s_tilt'])
        dir0 = np.array([-np.sin(t_), 0., np.cos(t_)])
        cols = [c for c in ALL_SCAT if c not in ('position', 'E')]
        saved = ss.scatter_continuous_classical
        if model != 'classical':
            cls = _JL if model == 'JL' else _EX
            cls._mat = m
            ss.scatter_continuous_classical = cls               # trajectory_cont_cl looks the name up at call time
        res = []
        try:
            for eid in ids:
                inj.start(eid)
                e = rel.electron(ip['E0'], ip['Emin'], np.zeros(3), dir0, {'el_output': ALL_EL, 'scat_output': cols})
                ss.trajectory_cont_cl(e, m, ip['maxScatt'], ip['thickness'], ip['elastic'], {})
                res.append((e.el_output.dict, e.scat_output.dict))
        finally:
            ss.scatter_continuous_classical = saved
        return flatten_traj(res, ['Ruth_vanilla'])

    ipc = base_input(material='Al', Emin=500., mode='cont', maxScatt=100000)
    for model in ('classical', 'JL', 'explicit'):
        np.savez_compressed(os.path.join(GOLD, 'traj_cont_Al_%s.npz' % model), ids=np.arange(6),
                            **run_cont(dict(ipc, Bethe_model=model), range(6), model))
    random.random = saved_random
    sc.scatter_discrete.det_type = orig_det
    sc.pickTable = sc_pickTable
    tot = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print('golden written: %d files, %.1f kB' % (len(os.listdir(GOLD)), tot / 1e3))


if __name__ == '__main__':
    main()
