"""Pin the oracle port (oracle/mcd_oracle.py) to the golden vectors produced by the reference's own
code (oracle/gen_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

import mcd_oracle as o

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
RTOL = 1e-12      # unit-factor association differs at the ulp level; Quinn's sqrt difference amplifies it ~2e3x


def load(name):
    return np.load(os.path.join(GOLD, name))


def test_material_and_formulas():
    F = json.load(open(os.path.join(GOLD, 'formulas.json')))
    for name, g in F.items():
        m = o.Material(name)
        assert m.atnd == pytest.approx(g['atnd'], rel=RTOL)
        assert m.fermi_e == pytest.approx(g['fermi_e'], rel=RTOL)
        assert m.plasmon_e == pytest.approx(g['plasmon_e'], rel=RTOL)
        for r in g['rows']:
            E = r['E']
            sig = o.sigmas_ds(E, m, 10.)
            np.testing.assert_allclose(sig, r['sigma'], rtol=RTOL, atol=0)
            assert o.mfp_from_sigma(sum(sig), m.atnd) == pytest.approx(r['mfp_total'], rel=RTOL)
            assert o.mfp_from_sigma(sig[0], m.atnd) == pytest.approx(r['mfp_ruth'], rel=RTOL)
            assert o.stopping_power('classical', E, m) == pytest.approx(r['sp_classical'], rel=RTOL)
            assert o.stopping_power('JL', E, m) == pytest.approx(r['sp_JL'], rel=RTOL)
            assert o.stopping_power('explicit', E, m) == pytest.approx(r['sp_explicit'], rel=RTOL)


def test_survey_kats():
    """SURVEY.md 8(c) probe-derived known answers (Si, 20 keV, Wc=10)."""
    m = o.Material('Si')
    sig = o.sigmas_ds(20000., m, 10.)
    np.testing.assert_allclose(sig, [8.395922089e-18, 1.297214282e-18, 1.021125579e-20, 1.181771096e-19,
                                     5.5376065e-19, 5.425130063e-18], rtol=2e-9)
    assert o.mfp_from_sigma(sum(sig), m.atnd) == pytest.approx(126.677387594071, rel=1e-9)


def test_pick_from_sigmas_bit_exact():
    g = load('select.npz')
    for row, u, pick in zip(g['sigma'], g['u'], g['pick']):
        sig = [float(x) for x in row[row >= 0]]
        assert o.pick_from_sigmas(sig, float(u)) == int(pick)


def test_angles_and_rotation():
    g = load('angles.npz')
    c2 = [o.rutherford_halfpol(float(E), int(Z), float(u)) for E, Z, u in zip(g['E'], g['Z'], g['u'])]
    np.testing.assert_allclose(c2, g['c2_ruth'], rtol=RTOL)
    cb = np.array([o.binary_coll(float(E), float(W), float(Ef)) for E, W, Ef in zip(g['E'], g['W'], g['Ef'])])
    assert np.array_equal(np.isnan(cb), np.isnan(g['c2_bin']))
    ok = ~np.isnan(cb)
    np.testing.assert_allclose(cb[ok], g['c2_bin'][ok], rtol=RTOL)
    for i in range(len(g['c2_in'])):
        d, y = o.update_direction(g['d_in'][i], g['y_in'][i], float(g['c2_in'][i]), float(g['hp_in'][i]))
        np.testing.assert_allclose(d, g['d_out'][i], rtol=0, atol=1e-14)
        np.testing.assert_allclose(y, g['y_out'][i], rtol=0, atol=1e-14)


@pytest.mark.parametrize('key', ['Si_15k_small', 'Al_500_small', 'Cu_10k_small'])
def test_tables_bit_exact(key):
    ax = json.load(open(os.path.join(GOLD, 'table_axes.json')))[key]
    g = load('tables_%s.npz' % key)
    ip = dict(ax['input'])
    tabs = o.gen_tables(ip, strict=True)
    for nm, T in (('moller', tabs['Moller']), ('gryz0', tabs['Gryz'][0])):
        assert np.array_equal(T.Ws, g[nm + '_Ws'])
        assert np.array_equal(T.Es, g[nm + '_Es'])
        assert T.table.dtype == np.float32
        assert np.array_equal(T.table, g[nm + '_table'])
    assert tabs['Gryz'][0].shell == str(g['gryz0_shell'])
    # pickTable lookups
    for nm, T in (('moller', tabs['Moller']), ('gryz0', tabs['Gryz'][0])):
        W = [o.pick_table(T, float(E), float(u))[0] for E, u in zip(g['lookup_E'], g['lookup_u'])]
        assert np.array_equal(np.asarray(W, dtype=np.float64), g[nm + '_lookup_W'])


@pytest.mark.parametrize('key', ['Si_10k_mid', 'Al_500_mid'])
def test_table_axes_and_rows_mid(key):
    ax = json.load(open(os.path.join(GOLD, 'table_axes.json')))[key]
    g = load('tablerows_%s.npz' % key)
    tabs = o.gen_tables(dict(ax['input']))
    assert [len(tabs['Moller'].Es), len(tabs['Moller'].Ws)] == ax['moller']
    assert [tabs['Gryz'][0].shell, len(tabs['Gryz'][0].Es), len(tabs['Gryz'][0].Ws)] == ax['gryz'][0]
    assert np.array_equal(tabs['Moller'].table[g['moller_rows_idx']], g['moller_rows'])
    assert np.array_equal(tabs['Gryz'][0].table[g['gryz0_rows_idx']], g['gryz0_rows'])


@pytest.mark.parametrize('key', ['C2_Si', 'C3_Cu', 'C4_Au'])
def test_table_axes_and_rows_default_tolerance(key):
    """BASELINE's DEFAULT tolerances (tol_E 1e-4, tol_W 1e-7): axes (SHA-1 of the float32 arrays) and eight rows per
    table bit-exact against the reference's own probTable (oracle/gen_golden_defaults.py), and pickTable on those
    rows."""
    import hashlib
    ax = json.load(open(os.path.join(GOLD, 'table_axes_default.json')))[key]
    g = load('tablerows_%s_default.npz' % key)
    ip = ax['input']
    mat = o.Material(ip['material'])
    sha = lambda a: hashlib.sha1(np.ascontiguousarray(a, np.float32).tobytes()).hexdigest()
    for nm, T in (('moller', o.ProbTable('Moller', mat, (ip['Emin'], ip['E0']), ip['tol_E'], ip['tol_W'], Wc=ip['Wc'])),
                  ('gryz0', o.ProbTable('Gryzinski', mat, (ip['Emin'], ip['E0']), ip['tol_E'], ip['tol_W'],
                                        shell=mat.name_s[-1]))):
        T.set_axes()
        a = ax[nm]
        assert [T.shell, len(T.Es), len(T.Ws)] == [a['shell'], a['nE'], a['nW']]
        assert sha(T.Es) == a['Es_sha1'] and sha(T.Ws) == a['Ws_sha1']
        idx = g[nm + '_rows_idx']
        rows, _ = T.compute_rows(T.Es[idx])
        assert np.array_equal(rows, g[nm + '_rows'])
        # pickTable on the stored rows (sparse table: only those rows exist)
        T.table = {int(i): rows[k] for k, i in enumerate(idx)}
        T.masked = {int(i): rows[k] == 0 for k, i in enumerate(idx)}
        W = [o.pick_table(T, float(E), float(u))[0] for E, u in zip(g[nm + '_lookup_E'], g[nm + '_lookup_u'])]
        assert np.array_equal(np.asarray(W, dtype=np.float64), g[nm + '_lookup_W'])


def _tables_from_golden(key, mat):
    g = load('tables_%s.npz' % key) if not key.endswith('.npz') else load(key)
    out = {}
    for nm, kind in (('moller', 'Moller'), ('gryz0', 'Gryz')):
        T = o.ProbTable.__new__(o.ProbTable)
        T.Ws, T.Es, T.table = g[nm + '_Ws'], g[nm + '_Es'], g[nm + '_table']
        T.masked = T.table == 0
        out[kind] = T if kind == 'Moller' else [T]
    return out


def _check_traj(g, els, scat_cols, atol=1e-9):
    n_ev = [len(e.scats) for e in els]
    assert n_ev == list(g['n_ev'])
    assert [len(e.els) for e in els] == list(g['n_els'])
    rows = [r for e in els for r in e.els]
    assert [o.OUTCOMES.index(r[0]) for r in rows] == list(g['outcome'])
    np.testing.assert_allclose([r[1] for r in rows], g['final_E'], rtol=1e-12)
    np.testing.assert_allclose(np.array([r[2] for r in rows]), g['final_dir'], atol=1e-11)
    np.testing.assert_allclose(np.array([r[3] for r in rows]), g['last_pos'], rtol=1e-10, atol=atol)
    ev = [r for e in els for r in e.scats]
    if 'type' in scat_cols:
        assert [r['channel'] for r in ev] == list(g['type'])
    for k in scat_cols:
        if k == 'type' or g[k].size == 0:
            continue
        mine = np.array([r[k] for r in ev], dtype=float)
        np.testing.assert_allclose(mine, g[k], rtol=1e-10, atol=atol, equal_nan=True)


@pytest.mark.parametrize('fname,tkey,kw', [
    ('traj_DS_Si.npz', 'Si_15k_small', {}),
    ('traj_DS_Cu.npz', 'Cu_10k_small', dict(material='Cu', Emin=10000.)),
    ('traj_DS_Si_long.npz', 'Si_15k_small', dict(maxScatt=5)),
])
def test_trajectories_DS(fname, tkey, kw):
    g = load(fname)
    ip = dict(E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, mode='DS', maxScatt=1000, Wc=10.)
    ip.update(kw)
    mat = o.Material(ip['material'])
    tabs = _tables_from_golden(tkey, mat)
    ids = g['ids']
    els = o.run_electrons(ip, int(ids[0]), len(ids), 12345, tabs, mat)
    _check_traj(g, els, ['type', 'position', 'pathl', 'E', 'E_loss', 'pol_angle', 'az_angle', 'energy', 'direction'])


def test_trajectories_DS_q4_double_record():
    """a 'bks' electron whose exit-event loss crosses Emin is recorded again as 'abs'."""
    g = load('traj_DS_Si_q4.npz')
    ip = dict(E0=20000., Emin=19985., material='Si', s_tilt=89.5, thickness=3e6, mode='DS', maxScatt=1000, Wc=10.)
    mat = o.Material('Si')
    tabs = _tables_from_golden('traj_DS_Si_q4.npz', mat)
    els = o.run_electrons(ip, 0, 300, 12345, tabs, mat)
    assert [len(e.scats) for e in els] == list(g['n_ev'])
    assert [len(e.els) for e in els] == list(g['n_els'])
    assert (g['n_els'] == 2).sum() > 0
    rows = [r for e in els for r in e.els]
    assert [o.OUTCOMES.index(r[0]) for r in rows] == list(g['outcome'])
    np.testing.assert_allclose([r[1] for r in rows], g['final_E'], rtol=1e-12)


@pytest.mark.parametrize('model', ['classical', 'JL', 'explicit'])
def test_trajectories_cont(model):
    g = load('traj_cont_Al_%s.npz' % model)
    ip = dict(E0=20000., Emin=500., material='Al', s_tilt=70., thickness=3e6, mode='cont', Bethe_model=model,
              maxScatt=100000, Wc=10.)
    els = o.run_electrons(ip, 0, 6, 12345)
    _check_traj(g, els, ['pathl', 'E_loss', 'pol_angle', 'az_angle', 'energy', 'direction'], atol=1e-8)


def test_elastic_models_formulas_and_reference_behaviour():
    """the alternative elastic cross-sections (crossSections.py:63-143) and what the reference does
    with each label (recorded by oracle/gen_golden_elastic.py)."""
    F = json.load(open(os.path.join(GOLD, 'elastic_formulas.json')))
    for r in F['rows']:
        E, Z = r['E'], r['Z']
        assert o.ruther_sigma(E, Z) == pytest.approx(r['vanilla'], rel=RTOL)
        assert o.ruther_sigma_wDefl(E, Z) == pytest.approx(r['vanilla_wDefl'], rel=RTOL)
        assert o.ruther_N_sigma(E, Z) == pytest.approx(r['nigram'], rel=RTOL)
        assert o.ruther_N_sigma_wDefl(E, Z) == pytest.approx(r['nigram_wDefl'], rel=RTOL)
        # continuous mode: the _wDefl branches are unreachable (scattering.py:781-789)
        assert o.elastic_sigma(E, Z, 'Ruth_nigram_wDefl', 'cont') == o.ruther_N_sigma(E, Z)
        assert o.elastic_sigma(E, Z, 'Ruth_vanilla_wDefl', 'cont') == o.ruther_sigma(E, Z)
    b = F['reference_behaviour']
    assert b['DS Ruth_vanilla_wDefl'].startswith('NameError')
    assert b['cont Ruth_nigram_wDefl'] == 'identical to Ruth_nigram'
    assert b['cont Ruth_vanilla_wDefl'] == 'identical to Ruth_vanilla'
    with pytest.raises(ValueError):
        o.elastic_sigma(2e4, 14, 'Mott')


@pytest.mark.parametrize('elastic', ['Ruth_nigram', 'Ruth_nigram_wDefl'])
def test_trajectories_DS_elastic_models(elastic):
    g = load('traj_DS_Si_%s.npz' % elastic[5:])
    ip = dict(E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, mode='DS', maxScatt=1000, Wc=10.,
              elastic=elastic)
    mat = o.Material('Si')
    tabs = _tables_from_golden('Si_15k_small', mat)
    ids = g['ids']
    els = o.run_electrons(ip, int(ids[0]), len(ids), 12345, tabs, mat)
    _check_traj(g, els, ['type', 'position', 'pathl', 'E', 'E_loss', 'pol_angle', 'az_angle', 'energy', 'direction'])


@pytest.mark.parametrize('elastic', ['Ruth_nigram', 'Ruth_nigram_wDefl'])
def test_trajectories_cont_elastic_models(elastic):
    g = load('traj_cont_Al_nigram.npz')                # the _wDefl label runs the same code in the reference
    ip = dict(E0=20000., Emin=500., material='Al', s_tilt=70., thickness=3e6, mode='cont', Bethe_model='classical',
              maxScatt=100000, Wc=10., elastic=elastic)
    els = o.run_electrons(ip, 0, 6, 12345)
    _check_traj(g, els, ['pathl', 'E_loss', 'pol_angle', 'az_angle', 'energy', 'direction'], atol=1e-8)
    assert els[0].scats[0]['type'] == elastic          # scattering.py:834


def test_trajectories_DS_thin_film_and_Au():
    """oracle/gen_golden_more.py: 'trsm' outcomes on a 9000 A film (electron.py:73-81) and Au with its 8 core shells."""
    g = load('traj_DS_Si_thin.npz')
    ip = dict(E0=20000., Emin=15000., material='Si', s_tilt=0., thickness=9000., mode='DS', maxScatt=1000, Wc=10.)
    mat = o.Material('Si')
    els = o.run_electrons(ip, int(g['ids'][0]), len(g['ids']), 12345, _tables_from_golden('Si_15k_small', mat), mat)
    cols = ['type', 'position', 'pathl', 'E', 'E_loss', 'pol_angle', 'az_angle', 'energy', 'direction']
    _check_traj(g, els, cols)
    assert (g['outcome'] == 2).sum() > 10 and (g['outcome'] == 0).sum() > 0        # transmitted and absorbed
    g = load('traj_DS_Au.npz')
    ip = dict(E0=30000., Emin=27000., material='Au', s_tilt=0., thickness=3e6, mode='DS', maxScatt=1000, Wc=10.)
    mat = o.Material('Au')
    els = o.run_electrons(ip, 0, 3, 12345, _tables_from_golden('traj_DS_Au.npz', mat), mat)
    _check_traj(g, els, cols)
    assert len(np.unique(g['type'])) >= 8                                           # most of the 11 channels occur


# ---- elastic = 'Mott' (extFunctions.py:21-153, scattering.py:96-122, :493, :588-591) on the synthetic ioffe-layout
# fixture tests/data/tables/mott/ioffe/Si.ioffe; goldens by the reference's own classes (oracle/gen_golden_mott.py) ----
MOTT_PATH = os.path.join(os.path.dirname(__file__), 'data', 'tables', 'mott') + os.sep


def _mott_oracle():
    return o.MottTable(14, 'Si', MOTT_PATH).read_ioffe().to_prob_table()


def test_mott_table_parse_cdf_and_picks():
    g = load('mott_Si.npz')
    t = _mott_oracle()
    for k in ('Es', 'sigmas', 'thetas', 'dCS', 'probTable'):
        assert np.array_equal(getattr(t, k), g[k]), k                      # bit-exact: same numpy operations
    assert (np.diff(t.Es) > 0).all() and t.probTable.shape == (56, 101)
    c2 = [o.pick_mott_table(t, float(E), float(u))[0] for E, u in zip(g['pick_E'], g['pick_u'])]
    assert np.array_equal(np.array(c2), g['pick_c2'])
    sg = [o.mott_sigma(t, float(E)) for E in g['pick_E']]
    assert np.array_equal(np.array(sg), g['pick_sigma'])
    # what the reference does with Mott in continuous mode is recorded, not assumed (the product refuses it)
    assert 'tableMott' in str(g['cont_behaviour'])
    with pytest.raises(ValueError):
        o.MottTable(13, 'Si', MOTT_PATH)                                     # Z check, extFunctions.py:58-60


def test_trajectories_DS_mott():
    g = load('traj_DS_Si_mott.npz')
    ip = dict(E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, mode='DS', maxScatt=1000, Wc=10.,
              elastic='Mott')
    mat = o.Material('Si')
    tabs = _tables_from_golden('Si_15k_small', mat)
    tabs['Mott'] = _mott_oracle()
    ids = g['ids']
    els = o.run_electrons(ip, int(ids[0]), len(ids), 12345, tabs, mat)
    _check_traj(g, els, ['type', 'position', 'pathl', 'E', 'E_loss', 'pol_angle', 'az_angle', 'energy', 'direction'])
    assert int(g['n_mott']) > 500 and els[0].scats[0]['type'] in ('Mott', 'Moller', 'Quinn', 'Gryz1s', 'Gryz2s', 'Gryz2p')
    assert sum(r['type'] == 'Mott' for e in els for r in e.scats) == int(g['n_mott'])


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10."""
    assert o.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert o.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert o.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
