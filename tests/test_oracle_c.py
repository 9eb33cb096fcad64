"""Pin the C restatement (oracle/mcd_oracle_c.c) event for event to the numpy/Python oracle, which is
itself pinned to the reference's golden vectors (tests/test_oracle_golden.py).  CPU only."""
import os

import numpy as np
import pytest

import mcd_oracle as o
import mcd_oracle_cbind as oc

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _golden_tabs(key):
    g = np.load(os.path.join(GOLD, 'tables_%s.npz' % key))
    out = {}
    for nm, kind in (('moller', 'Moller'), ('gryz0', 'Gryz')):
        T = o.ProbTable.__new__(o.ProbTable)
        T.Ws, T.Es, T.table = g[nm + '_Ws'], g[nm + '_Es'], g[nm + '_table']
        T.masked = T.table == 0
        out[kind] = T if kind == 'Moller' else [T]
    return out


@pytest.mark.parametrize('name,key,Emin,tilt', [('Si', 'Si_15k_small', 15000., 70.), ('Cu', 'Cu_10k_small', 10000., 70.),
                                                ('Al', 'Al_500_small', 500., 0.)])
def test_c_port_matches_python_port_DS(name, key, Emin, tilt):
    ip = dict(E0=20000., Emin=Emin, material=name, s_tilt=tilt, thickness=3e6, mode='DS', maxScatt=1000, Wc=10.)
    tabs = _golden_tabs(key)
    n = 24
    py = o.run_electrons(ip, 50, n, 999, tabs)
    total_ev = sum(len(e.scats) for e in py)
    c = oc.run(ip, 50, n, 999, tabs, event_rows=total_ev)
    assert list(c['n_ev']) == [len(e.scats) for e in py]
    assert list(c['outcome']) == [o.OUTCOMES.index(e.els[0][0]) for e in py]
    assert list(c['n_rows']) == [len(e.els) for e in py]
    np.testing.assert_allclose(c['final_E'], [e.els[0][1] for e in py], rtol=1e-12)
    np.testing.assert_allclose(c['final_dir'], [e.els[0][2] for e in py], atol=1e-11)
    np.testing.assert_allclose(c['last_pos'], [e.els[0][3] for e in py], rtol=1e-10, atol=1e-8)
    ev = [r for e in py for r in e.scats]
    assert list(c['events'][:, 0].astype(int)) == [r['channel'] for r in ev]
    np.testing.assert_allclose(c['events'][:, 1:4], [r['position'] for r in ev], rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(c['events'][:, 4], [r['pathl'] for r in ev], rtol=1e-12)
    np.testing.assert_allclose(c['events'][:, 5], [r['E_loss'] for r in ev], rtol=1e-12)
    np.testing.assert_allclose(c['events'][:, 7], [r['energy'] for r in ev], rtol=1e-12)


@pytest.mark.parametrize('model', ['classical', 'JL', 'explicit'])
def test_c_port_matches_python_port_cont(model):
    ip = dict(E0=20000., Emin=500., material='Al', s_tilt=70., thickness=3e6, mode='cont', Bethe_model=model,
              maxScatt=100000, Wc=10.)
    py = o.run_electrons(ip, 0, 12, 12345)
    c = oc.run(ip, 0, 12, 12345)
    assert list(c['n_ev']) == [len(e.scats) for e in py]
    assert list(c['outcome']) == [o.OUTCOMES.index(e.els[0][0]) for e in py]
    np.testing.assert_allclose(c['final_E'], [e.els[0][1] for e in py], rtol=1e-12)


@pytest.mark.parametrize('elastic', o.ELASTIC_MODELS[1:])
def test_c_port_matches_python_port_elastic_models(elastic):
    """DS and continuous runs with the alternative elastic cross-sections."""
    ip = dict(E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, mode='DS', maxScatt=1000, Wc=10.,
              elastic=elastic)
    tabs = _golden_tabs('Si_15k_small')
    py = o.run_electrons(ip, 7, 16, 4242, tabs)
    c = oc.run(ip, 7, 16, 4242, tabs)
    assert list(c['n_ev']) == [len(e.scats) for e in py]
    assert list(c['outcome']) == [o.OUTCOMES.index(e.els[0][0]) for e in py]
    np.testing.assert_allclose(c['final_E'], [e.els[0][1] for e in py], rtol=1e-12)
    np.testing.assert_allclose(c['last_pos'], [e.els[0][3] for e in py], rtol=1e-10, atol=1e-8)
    ipc = dict(ip, material='Al', Emin=500., mode='cont', Bethe_model='JL', maxScatt=100000)
    py = o.run_electrons(ipc, 0, 8, 4242)
    c = oc.run(ipc, 0, 8, 4242)
    assert list(c['n_ev']) == [len(e.scats) for e in py]
    np.testing.assert_allclose(c['final_E'], [e.els[0][1] for e in py], rtol=1e-12)


def test_c_port_thin_film_and_Au_goldens():
    """The C port against the reference's own rows for the thin film ('trsm') and Au (8 shells) cases."""
    g = np.load(os.path.join(GOLD, 'traj_DS_Si_thin.npz'))
    ip = dict(E0=20000., Emin=15000., material='Si', s_tilt=0., thickness=9000., mode='DS', maxScatt=1000, Wc=10.)
    c = oc.run(ip, int(g['ids'][0]), len(g['ids']), 12345, _golden_tabs('Si_15k_small'))
    assert np.array_equal(c['outcome'], g['outcome']) and np.array_equal(c['n_ev'], g['n_ev'])
    np.testing.assert_allclose(c['final_E'], g['final_E'], rtol=1e-12)
    np.testing.assert_allclose(c['last_pos'], g['last_pos'], rtol=1e-10, atol=1e-8)
    g = np.load(os.path.join(GOLD, 'traj_DS_Au.npz'))
    T = {}
    for nm, kind in (('moller', 'Moller'), ('gryz0', 'Gryz')):
        t = o.ProbTable.__new__(o.ProbTable)
        t.Ws, t.Es, t.table = g[nm + '_Ws'], g[nm + '_Es'], g[nm + '_table']
        T[kind] = t if kind == 'Moller' else [t]
    ip = dict(E0=30000., Emin=27000., material='Au', s_tilt=0., thickness=3e6, mode='DS', maxScatt=1000, Wc=10.)
    c = oc.run(ip, 0, 3, 12345, T, event_rows=int(g['n_ev'].sum()))
    assert np.array_equal(c['outcome'], g['outcome']) and np.array_equal(c['n_ev'], g['n_ev'])
    assert np.array_equal(c['events'][:, 0].astype(int), g['type'])
    np.testing.assert_allclose(c['events'][:, 7], g['energy'], rtol=1e-12)


def test_c_port_threads_and_q4():
    g = np.load(os.path.join(GOLD, 'traj_DS_Si_q4.npz'))
    T = {}
    for nm, kind in (('moller', 'Moller'), ('gryz0', 'Gryz')):
        t = o.ProbTable.__new__(o.ProbTable)
        t.Ws, t.Es, t.table = g[nm + '_Ws'], g[nm + '_Es'], g[nm + '_table']
        T[kind] = t if kind == 'Moller' else [t]
    ip = dict(E0=20000., Emin=19985., material='Si', s_tilt=89.5, thickness=3e6, mode='DS', maxScatt=1000, Wc=10.)
    a = oc.run(ip, 0, 300, 12345, T, nthreads=1)
    b = oc.run(ip, 0, 300, 12345, T, nthreads=4)
    assert np.array_equal(a['outcome'], b['outcome']) and np.array_equal(a['final_E'], b['final_E'])
    # the reference's own rows (golden): first outcome per electron and the double records
    first = np.concatenate([[0], np.cumsum(g['n_els'])])[:-1]
    assert np.array_equal(a['outcome'], g['outcome'][first])
    assert np.array_equal(a['n_rows'], g['n_els'])
    assert np.array_equal(a['n_ev'], g['n_ev'])


# ---- the reference's OWN random numbers: tests/golden/stats_*.npz are samples of the unmodified reference run with
# `random.seed(SEED + id)` per electron (oracle/gen_golden_stats.py), no Philox injection anywhere.  The C port (the
# statistical oracle of the GPU tests) must draw from the same distributions. ---------------------------------------
STAT_CASES = [('DS_Si', {}, 'Si_15k_small'), ('DS_Cu', dict(material='Cu', Emin=10000.), 'Cu_10k_small'),
              ('cont_Al', dict(material='Al', Emin=500., mode='cont', Bethe_model='classical', maxScatt=100000), None),
              ('DS_Si_thin', dict(thickness=9000., s_tilt=0.), 'Si_15k_small')]       # 'trsm' outcomes


def reference_sample_checks(sample, ref, n_sample):
    """Backscatter coefficient within 3 sigma, two-sample KS (p > 0.001: fifteen tests in the family) on final
    energy / exit polar angle / azimuth / last depth of the backscattered and last depth of the absorbed electrons,
    mean trajectory length within 4 standard errors.  `sample`, `ref`: dicts with outcome, final_E, final_dir,
    last_pos, n_ev."""
    from scipy import stats
    n_ref = len(ref['outcome'])
    eg, er = (sample['outcome'] == 1).mean(), (ref['outcome'] == 1).mean()
    sd = np.sqrt(eg * (1 - eg) / n_sample + er * (1 - er) / n_ref)
    assert abs(eg - er) < 3 * sd, (eg, er, sd)
    for code in (0, 2, 4):                                   # abs, trsm, long fractions too
        fg, fr = (sample['outcome'] == code).mean(), (ref['outcome'] == code).mean()
        assert abs(fg - fr) <= 3 * np.sqrt(fg * (1 - fg) / n_sample + fr * (1 - fr) / n_ref) + 1e-12, (code, fg, fr)
    b, rb = sample['outcome'] == 1, ref['outcome'] == 1
    rng = np.random.default_rng(0)
    jit = lambda x: np.asarray(x, np.float64) + rng.uniform(-0.05, 0.05, len(x))      # breaks the ties of the energy comb
    f64 = lambda x: np.asarray(x, np.float64)
    pairs = {
        'final_E': (jit(sample['final_E'][b]), jit(ref['final_E'][rb])),
        'exit_polar': (np.arccos(np.clip(-f64(sample['final_dir'][b, 2]), -1, 1)), np.arccos(np.clip(-f64(ref['final_dir'][rb, 2]), -1, 1))),
        'exit_azimuth': (np.arctan2(f64(sample['final_dir'][b, 1]), f64(sample['final_dir'][b, 0])),
                         np.arctan2(f64(ref['final_dir'][rb, 1]), f64(ref['final_dir'][rb, 0]))),
        'depth_last_z': (f64(sample['last_pos'][b, 2]), f64(ref['last_pos'][rb, 2])),
        'abs_last_z': (f64(sample['last_pos'][sample['outcome'] == 0, 2]), f64(ref['last_pos'][ref['outcome'] == 0, 2])),
        'trsm_final_E': (jit(sample['final_E'][sample['outcome'] == 2]), jit(ref['final_E'][ref['outcome'] == 2])),
        'trsm_exit_polar': (np.arccos(np.clip(f64(sample['final_dir'][sample['outcome'] == 2, 2]), -1, 1)),
                            np.arccos(np.clip(f64(ref['final_dir'][ref['outcome'] == 2, 2]), -1, 1))),
    }
    ps = {k: float(stats.ks_2samp(a, c).pvalue) for k, (a, c) in pairs.items() if len(a) >= 50 and len(c) >= 50}
    se = ref['n_ev'].std() / np.sqrt(n_ref)
    print('eta %.4f vs reference %.4f (sd %.4f); KS p %s; mean events %.2f vs %.2f (se %.2f)'
          % (eg, er, sd, {k: '%.3g' % v for k, v in ps.items()}, sample['n_ev'].mean(), ref['n_ev'].mean(), se))
    assert all(v > 1e-3 for v in ps.values()), ps
    assert abs(sample['n_ev'].mean() - ref['n_ev'].mean()) < 4 * se
    return ps


@pytest.mark.parametrize('tag,kw,tkey', STAT_CASES)
def test_c_port_draws_from_the_reference_distributions(tag, kw, tkey):
    ref = dict(np.load(os.path.join(GOLD, 'stats_%s.npz' % tag)))
    ip = dict(E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, mode='DS', maxScatt=1000, Wc=10.)
    ip.update(kw)
    n = 60000
    c = oc.run(ip, 0, n, 31337, _golden_tabs(tkey) if tkey else None, nthreads=os.cpu_count() or 1)
    assert len(ref['outcome']) >= 4000 and int(ref['double_records'].max()) <= 1
    reference_sample_checks(c, ref, n)


def _mott_tabs():
    tabs = _golden_tabs('Si_15k_small')
    tabs['Mott'] = o.MottTable(14, 'Si', os.path.join(os.path.dirname(__file__), 'data', 'tables', 'mott') + os.sep).read_ioffe().to_prob_table()
    return tabs


def test_c_port_mott_matches_python_port_and_the_reference_distribution():
    """elastic = 'Mott' (synthetic ioffe-layout table): the C port follows the Python port event for event (which
    follows the reference's own golden trajectories, test_oracle_golden.py) and draws from the distributions of the
    reference's own-RNG sample."""
    ip = dict(E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, mode='DS', maxScatt=1000, Wc=10., elastic='Mott')
    tabs = _mott_tabs()
    py = o.run_electrons(ip, 50, 24, 999, tabs)
    ev = [r for e in py for r in e.scats]
    c = oc.run(ip, 50, 24, 999, tabs, event_rows=len(ev))
    assert list(c['n_ev']) == [len(e.scats) for e in py]
    assert list(c['outcome']) == [o.OUTCOMES.index(e.els[0][0]) for e in py]
    assert list(c['events'][:, 0].astype(int)) == [r['channel'] for r in ev]
    np.testing.assert_allclose(c['events'][:, 6], [r['pol_angle'] for r in ev], rtol=1e-14)
    np.testing.assert_allclose(c['final_E'], [e.els[0][1] for e in py], rtol=1e-12)
    ref = dict(np.load(os.path.join(GOLD, 'stats_DS_Si_mott.npz')))
    n = 60000
    reference_sample_checks(oc.run(ip, 0, n, 31337, tabs, nthreads=os.cpu_count() or 1), ref, n)
    with pytest.raises(ValueError):
        oc.run(dict(ip, mode='cont'), 0, 4, 1, tabs)
