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
