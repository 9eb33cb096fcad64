"""GPU: the doScatter-style command line end to end (input.file -> tables -> run -> store) and the drop-in
workers with the reference's signatures."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('fmt', ['parquet', 'hdf5'])
def test_cli_writes_reference_layout(tmp_path, fmt):
    inp = tmp_path / 'input.file'
    inp.write_text('\n'.join([
        'num_el : 3000', 'E0 : 20000 # eV', 'Emin : 15000', 'material : Si', 's_tilt : 70', 'thickness : 3000000',
        'elastic : Ruth_vanilla', 'mode : DS', 'Bethe_model : classical', 'diffMFP : no', 'gen_tables : yes',
        'tol_E : 1e-2', 'tol_W : 1e-3', 'maxScatt: 1000', 'Wc : 10',
        'electron_output : outcome, final_E, final_dir, last_pos', 'scatter_output : type, position, energy']))
    out = str(tmp_path / ('run.parquet' if fmt == 'parquet' else 'run.h5'))
    r = subprocess.run([sys.executable, '-m', 'mcdiscrete_b200', '-i', str(inp), '-o', out, '--format', fmt,
                        '--batch', '1000'], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    if fmt == 'hdf5':                  # the reference's store layout (keys input / els / scats, doScatter.py:63-67)
        assert open(out, 'rb').read(8) == b'\x89HDF\r\n\x1a\n'
    from mcdiscrete_b200.store import read_store
    back = read_store(out)
    assert len(back['els']) == 3000
    assert list(back['els'].columns) == ['outcome', 'final_E', 'dx', 'dy', 'dz', 'last_x', 'last_y', 'last_z']
    assert list(back['scats'].columns) == ['type', 'energy', 'x', 'y', 'z']
    assert set(back['els']['outcome']) <= {'abs', 'bks', 'trsm', 'long'}
    assert set(back['scats']['type']) <= {'Ruth', 'Moller', 'Gryz1s', 'Gryz2s', 'Gryz2p', 'Quinn'}
    first = back['scats'][(back['scats'][['x', 'y', 'z']] == 0).all(axis=1)]
    assert len(first) == 3000                                   # every trajectory starts at the origin
    st = json.load(open(out + '.stats.json'))
    assert sum(st['histograms']['energy']['counts']) == int((back['els']['outcome'] == 'bks').sum())


def test_workers_with_reference_signatures():
    from mcdiscrete_b200 import material, genTables, multiTraj_DS, multiTraj_cont, zipDict
    ip = dict(num_el=8, E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, elastic='Ruth_vanilla',
              mode='DS', Bethe_model='classical', diffMFP=False, gen_tables=True, tol_E=1e-2, tol_W=1e-3,
              maxScatt=1000, Wc=10., electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'],
              scatter_output=['type', 'position', 'energy'])
    what = {'el_output': ip['electron_output'], 'scat_output': ip['scatter_output']}
    mat = material(ip['material'])
    tables = genTables(ip)
    res = multiTraj_DS(ip, 8, mat, tables, what)                 # multiScatter.py:66 signature
    assert len(res) == 8 and set(res[0]) == {'els', 'scats'}
    r0 = res[0]
    assert list(r0['els']) == ['outcome', 'final_E', 'final_dir', 'last_pos']
    assert set(r0['els']['final_dir']) == {'dx', 'dy', 'dz'} and set(r0['els']['last_pos']) == {'last_x', 'last_y', 'last_z'}
    n = len(r0['scats']['type'])
    assert n == len(r0['scats']['energy']) == len(r0['scats']['position']['x']) >= 1
    assert r0['scats']['position']['x'][0] == 0.0 and r0['scats']['type'][0] in ('Ruth', 'Moller', 'Quinn', 'Gryz1s', 'Gryz2s', 'Gryz2p')
    # the listener's merge (doScatter.py:108-110) works on the returned structure
    merged = {}
    for r in res:
        merged = zipDict(merged, r['scats'])
    assert list(merged) == ['type', 'energy', 'x', 'y', 'z']
    again = multiTraj_DS(ip, 8, mat, tables, what)               # fresh electrons on every call
    assert [r['els']['final_E'] for r in again] != [r['els']['final_E'] for r in res]
    ipc = dict(ip, mode='cont', Emin=500., Bethe_model='JL', scatter_output=['type', 'pathl', 'E_loss', 'energy'])
    whatc = {'el_output': ip['electron_output'], 'scat_output': ipc['scatter_output']}
    resc = multiTraj_cont(ipc, 4, material('Si'), whatc, {})      # multiScatter.py:171 argument order
    assert len(resc) == 4 and resc[0]['scats']['type'][0] == 'Ruth_vanilla'
    assert len(resc[0]['scats']['pathl']) == len(resc[0]['scats']['energy'])


def test_error_behaviour_and_legacy_wrappers():
    """Status codes instead of crashes (include/mcd.h conventions) and the legacy process targets."""
    import queue
    from mcdiscrete_b200 import _lib
    from mcdiscrete_b200.engine import Engine
    from mcdiscrete_b200 import genTables, scatterMultiEl_DS, scatterMultiEl_cont
    ip = dict(num_el=6, E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, elastic='Ruth_vanilla',
              mode='DS', Bethe_model='classical', diffMFP=False, gen_tables=True, tol_E=1e-2, tol_W=1e-3,
              maxScatt=1000, Wc=10., electron_output=['outcome', 'final_E'], scatter_output=['type', 'energy'])
    eng = Engine(ip, device=0)
    eng.set_capacity(10, 1000)
    with pytest.raises(_lib.MCDError) as ei:                      # DS run without tables
        eng.launch(0, 10)
    assert ei.value.code == -4 and 'table' in str(ei.value)
    with pytest.raises(_lib.MCDError) as ei:                      # more electrons than the capacity
        genTables(ip, engine=eng)
        eng.launch(0, 11)
    assert ei.value.code == -1
    # too small a scats pool: explicit capacity -> error with the dropped-row count; automatic capacity regrows
    # (the pool always holds one spare chunk per resident warp, ~4.8e6 rows: overflow needs more rows than that)
    with pytest.raises(_lib.MCDError) as ei:
        eng.run(count=40000, first_id=0, max_scat_rows=1000)
    assert ei.value.code == -5
    res = eng.run(count=40000, first_id=0)
    assert res.counters['scat_rows_dropped'] == 0 and len(res.scats) == res.counters['events']
    # the same overflow through the pipelined columnar path, where the regrouping is already enqueued behind the run
    # when the dropped rows are noticed: clean error, nothing written past the columns, context usable afterwards
    eng2 = Engine(ip, device=0)
    genTables(ip, engine=eng2)
    with pytest.raises(_lib.MCDError) as ei:
        for _ in eng2.run_batches(80000, 40000, first_id=0, max_scat_rows=1000, ordered='columns'):
            pass
    assert ei.value.code == -5
    eng2.close()
    tot = sum(len(r.scats) for r in eng.run_batches(40000, 10000, first_id=0, ordered='columns'))
    assert tot == res.counters['events']
    with pytest.raises(ValueError):
        Engine(dict(ip, scatter_output=['nonsense']), device=0)
    with pytest.raises(NotImplementedError):
        Engine(dict(ip, elastic='Browning'), device=0)            # not a model of the reference's code
    with pytest.raises(_lib.MCDError):                            # Emin below the plasmon energy (quinn_sigma :272-281)
        Engine(dict(ip, Emin=10.), device=0)
    eng.close()
    # legacy targets (multiScatter.py:15-61, 118-168)
    tables = genTables(ip)
    what = {'el_output': ip['electron_output'], 'scat_output': ip['scatter_output']}
    out = {}
    scatterMultiEl_DS(ip, tables, what, out, 0)
    assert set(out) == {'els', 'scats'} and len(out['els']['outcome']) >= 1
    q = {'electrons': queue.Queue(), 'scatterings': queue.Queue()}
    ipc = dict(ip, mode='cont', Emin=500., scatter_output=['type', 'energy'])
    scatterMultiEl_cont(ipc, {'el_output': ip['electron_output'], 'scat_output': ipc['scatter_output']}, q, 5, 0)
    import pickle
    els = pickle.loads(q['electrons'].get())
    assert len(els['outcome']) == 5 and len(pickle.loads(q['scatterings'].get())['energy']) > 5


def test_raw_ctypes_sequence_of_INTEGRATION_md():
    """The bare C-ABI call sequence documented in INTEGRATION.md section 2 (no Engine wrapper)."""
    import ctypes as C
    from mcdiscrete_b200 import _lib, material
    from mcdiscrete_b200.probTables import probTable
    lib = C.CDLL(_lib.LIB_PATH)
    lib.mcd_last_error.restype = C.c_char_p

    def check(rc, ctx=None):
        assert rc == 0, lib.mcd_last_error(ctx).decode()
    inputPar = dict(E0=20000., Emin=15000., s_tilt=70., thickness=3e6, Wc=10., maxScatt=1000, num_el=5000, material='Si')
    mat = material(inputPar['material'])
    m = _lib.Material(mat.params['Z'], len(mat.params['name_s']), mat.params['n_val'], mat.params['E_val'],
                      mat.atnd, mat.fermi_e, mat.plasmon_e, mat.params['bethe_k'])
    for i, s in enumerate(mat.params['name_s']):
        m.shell_E[i], m.shell_n[i] = mat.params['Es'][s], mat.params['ns'][s]
    p = _lib.Params(E0=inputPar['E0'], Emin=inputPar['Emin'], s_tilt_deg=inputPar['s_tilt'], thickness=inputPar['thickness'],
                    Wc=inputPar['Wc'], maxScatt=inputPar['maxScatt'], mode=0, el_mask=0xF, scat_mask=0x83, flags=1, seed=12345)
    ctx = C.c_void_p()
    check(lib.mcd_create(C.byref(p), C.byref(m), 0, C.byref(ctx)))
    T = probTable('Moller', mat.params['name_val'], (inputPar['Emin'], inputPar['E0']), 1e-2, 1e-3, mat, Wc=inputPar['Wc'])
    T.set_Ws(); T.set_Es()
    G = probTable('Gryzinski', mat.params['name_s'][-1], (inputPar['Emin'], inputPar['E0']), 1e-2, 1e-3, mat)
    G.set_Ws(); G.set_Es()
    f = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(C.POINTER(C.c_float))
    check(lib.mcd_table_build(ctx, 0, f(T.Es), C.c_uint32(len(T.Es)), f(T.Ws), C.c_uint32(len(T.Ws)), C.c_double(0), C.c_double(0), None), ctx)
    check(lib.mcd_table_build(ctx, 1, f(G.Es), C.c_uint32(len(G.Es)), f(G.Ws), C.c_uint32(len(G.Ws)), C.c_double(G.Wmin),
                              C.c_double(G.numEl), None), ctx)
    n = inputPar['num_el']
    check(lib.mcd_set_capacity(ctx, C.c_uint64(n), C.c_uint64(n * 512)), ctx)
    check(lib.mcd_run(ctx, C.c_uint64(0), C.c_uint64(n), None), ctx)
    check(lib.mcd_sync(ctx, None), ctx)
    els = np.empty(n, dtype=_lib.EL_ROW_DTYPE)
    check(lib.mcd_drain_els(ctx, els.ctypes.data_as(C.c_void_p), C.c_uint64(0), C.c_uint64(n), None), ctx)
    check(lib.mcd_sync(ctx, None), ctx)
    outcome = np.array(['abs', 'bks', 'trsm', 'trdf', 'long'])[els['code'] & 0xff]
    assert set(outcome) <= {'abs', 'bks'} and 0.3 < (outcome == 'bks').mean() < 0.7
    n_els, n_sc = C.c_uint64(), C.c_uint64()
    check(lib.mcd_row_counts(ctx, C.byref(n_els), C.byref(n_sc)), ctx)
    assert n_els.value == n and n_sc.value > n
    lib.mcd_destroy(ctx)


def test_edge_cases_empty_and_degenerate_runs():
    from mcdiscrete_b200 import material, genTables, multiTraj_DS
    from mcdiscrete_b200.engine import Engine
    ip = dict(num_el=0, E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, elastic='Ruth_vanilla',
              mode='DS', Bethe_model='classical', diffMFP=False, gen_tables=True, tol_E=1e-2, tol_W=1e-3,
              maxScatt=1000, Wc=10., electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'],
              scatter_output=['type', 'energy'])
    what = {'el_output': ip['electron_output'], 'scat_output': ip['scatter_output']}
    tables = genTables(ip)
    assert multiTraj_DS(ip, 0, material('Si'), tables, what) == []           # no electrons: no rows, no error
    eng = Engine(dict(ip, maxScatt=0), device=0)                             # every electron is 'long' at once
    genTables(ip, engine=eng)
    r = eng.run(count=1000, first_id=0)
    assert (r.outcome == 4).all() and r.counters['events'] == 0 and len(r.scats) == 0
    assert (r.els['final_E'] == 20000.).all() and (r.els['last_z'] == 0).all()
    eng.close()
    eng = Engine(dict(ip, maxScatt=1), device=0)                             # exactly one event each
    genTables(ip, engine=eng)
    r = eng.run(count=1000, first_id=0, ordered=True)
    assert r.counters['events'] == 1000 and len(r.scats) == 1000
    assert np.array_equal(r.scats['el'], np.arange(1000)) and (r.scats['ev'] == 0).all()
    assert set(np.unique(r.outcome)) <= {0, 1, 4}
    eng.close()
    # one electron, very large id (64-bit Philox counter)
    eng = Engine(ip, device=0)
    genTables(ip, engine=eng)
    a = eng.run(count=1, first_id=2 ** 45 + 12345)
    b = eng.run(count=3, first_id=2 ** 45 + 12344)
    assert a.els[0] == b.els[1]
    eng.close()


def test_top_level_run():
    """mcdiscrete_b200.run(inputPar): parameters -> columnar results in one call."""
    import numpy as np
    import mcdiscrete_b200 as mcd
    ip = dict(num_el=2000, E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, elastic='Ruth_vanilla',
              mode='DS', Bethe_model='classical', diffMFP=False, gen_tables=True, tol_E=1e-2, tol_W=1e-3,
              maxScatt=1000, Wc=10., electron_output=['outcome', 'final_E'], scatter_output=['type', 'energy'])
    a = mcd.run(ip, seed=7, device=0)
    b = mcd.run(ip, seed=7, device=0, ordered=True)
    assert a.count == 2000 and len(a.els) == 2000 and a.counters['electrons_done'] == 2000
    assert a.scats.offsets[-1] == len(a.scats) == a.counters['events']
    assert np.array_equal(a.els, b.els) and np.array_equal(a.scats['energy'], b.scats['energy'])
    c = mcd.run(dict(ip, mode='cont', Emin=2000.), seed=7, device=0, count=500)
    assert c.count == 500 and set(np.unique(c.outcome)) <= {0, 1, 2, 4}


def test_workers_do_not_reuse_a_context_across_changed_parameters():
    """Two worker calls in one process that differ only in `elastic` or `thickness` (fields the round-1 engine cache
    ignored) must run with their own parameters -- also when the tables handed in were built on the first engine."""
    from mcdiscrete_b200 import material, genTables, multiTraj_DS
    from mcdiscrete_b200.engine import Engine
    ip = dict(num_el=4000, E0=20000., Emin=15000., material='Si', s_tilt=0., thickness=3e6, elastic='Ruth_vanilla',
              mode='DS', Bethe_model='classical', diffMFP=False, gen_tables=True, tol_E=1e-2, tol_W=1e-3,
              maxScatt=1000, Wc=10., electron_output=['outcome', 'final_E'], scatter_output=[], seed=99)
    what = {'el_output': ip['electron_output'], 'scat_output': []}
    mat = material('Si')
    tables = genTables(ip)                                      # built on an engine created from `ip`
    n = 4000

    def fractions(res):
        o = [r['els']['outcome'][0] for r in res]
        return {k: o.count(k) / len(o) for k in ('abs', 'bks', 'trsm')}
    thick = fractions(multiTraj_DS(ip, n, mat, tables, what))
    thin = fractions(multiTraj_DS(dict(ip, thickness=1500.), n, mat, tables, what))
    assert thick['trsm'] == 0 and thin['trsm'] > 0.5             # a 150 nm film transmits most 20 keV electrons
    nig = fractions(multiTraj_DS(dict(ip, elastic='Ruth_nigram', s_tilt=70.), n, mat, tables, what))
    van = fractions(multiTraj_DS(dict(ip, s_tilt=70.), n, mat, tables, what))
    # reference numbers for the two models from dedicated engines on the same ids
    want = {}
    for el in ('Ruth_nigram', 'Ruth_vanilla'):
        e = Engine(dict(ip, elastic=el, s_tilt=70.), device=0)
        genTables(dict(ip, elastic=el, s_tilt=70.), engine=e)
        want[el] = float((e.run(count=n, first_id=0).outcome == 1).mean())
        e.close()
    assert abs(nig['bks'] - want['Ruth_nigram']) < 1e-12 and abs(van['bks'] - want['Ruth_vanilla']) < 1e-12
    assert abs(want['Ruth_nigram'] - want['Ruth_vanilla']) > 0.01


def test_non_uniform_energy_axis_is_bisected():
    """mcd_table_upload accepts any strictly increasing Es; an axis that is not the uniform grid the row guess assumes
    (or is finer than float32 resolves) must be searched by bisection: rows identical to bisect_left."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import mcd_oracle as o
    from mcdiscrete_b200 import _lib
    from mcdiscrete_b200.engine import Engine
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'tables_Si_15k_small.npz'))
    ip = dict(E0=20000., Emin=15000., material='Si', s_tilt=70., maxScatt=1000, mode='DS', scatter_output=[])
    eng = Engine(ip, device=0)
    rng = np.random.default_rng(5)
    nE = len(g['moller_Es'])
    for Es in (np.sort(rng.uniform(15000., 20000., nE)).astype(np.float32),                  # irregular
               np.geomspace(15000., 20000., nE).astype(np.float32),                          # geometric
               (15000. + 5000. * np.linspace(0, 1, nE) ** 2).astype(np.float32)):            # quadratic
        Es[0], Es[-1] = 15000., 20000.
        assert (np.diff(Es) > 0).all()
        eng.table_upload(0, Es, g['moller_Ws'], g['moller_table'])
        E = np.concatenate([rng.uniform(15000.1, 20000., 4000), Es[1:], np.nextafter(Es[1:], np.float32(0)),
                            np.nextafter(Es[:-1], np.float32(1e9))]).astype(np.float32)
        u = np.clip(rng.random(len(E)), 2.0 ** -24, 1 - 2.0 ** -24).astype(np.float32)
        W, iE, jW = eng.test_lookup(0, E, u)
        assert np.array_equal(iE, np.searchsorted(Es, E, side='left'))
    with pytest.raises(_lib.MCDError) as ei:                    # not increasing: rejected
        bad = g['moller_Es'].copy(); bad[3] = bad[2]
        eng.table_upload(0, bad, g['moller_Ws'], g['moller_table'])
    assert ei.value.code == -1
    eng.close()


def test_mott_through_the_worker_and_the_command_line(tmp_path):
    """elastic: Mott end to end: genTables reads the ioffe-layout table and uploads it, multiTraj_DS labels the
    elastic channel 'Mott' (scattering.py:589) -- on the engine the tables were built on and on a fresh engine made
    for other output columns -- and the doScatter-style command line writes the label into the store."""
    from mcdiscrete_b200 import material, genTables, multiTraj_DS
    from mcdiscrete_b200.store import read_store
    mott_path = os.path.join(os.path.dirname(__file__), 'data', 'tables', 'mott') + os.sep
    ip = dict(num_el=40, E0=20000., Emin=15000., material='Si', s_tilt=70., thickness=3e6, elastic='Mott',
              mode='DS', Bethe_model='classical', diffMFP=False, gen_tables=True, tol_E=1e-2, tol_W=1e-3,
              maxScatt=1000, Wc=10., electron_output=['outcome', 'final_E'], scatter_output=['type', 'pol_angle'],
              mott_path=mott_path)
    tables = genTables(ip)
    assert set(tables) == {'Moller', 'Gryz', 'Mott'} and tables['Mott'].probTable.shape == (56, 101)
    what = {'el_output': ip['electron_output'], 'scat_output': ip['scatter_output']}
    res = multiTraj_DS(ip, 40, material('Si'), tables, what)
    types = [t for r in res for t in r['scats']['type']]
    assert 'Mott' in types and 'Ruth' not in types and set(types) <= {'Mott', 'Moller', 'Quinn', 'Gryz1s', 'Gryz2s', 'Gryz2p'}
    pol = np.array([c for r in res for t, c in zip(r['scats']['type'], r['scats']['pol_angle']) if t == 'Mott'])
    grid = np.cos(np.radians(tables['Mott'].thetas * 0.5)) ** 2
    assert np.abs(pol[:, None] - grid[None, :]).min(axis=1).max() < 1e-6        # only angles of the table's grid
    what2 = {'el_output': ['outcome'], 'scat_output': ['type', 'energy']}       # another context: tables are moved over
    res2 = multiTraj_DS(ip, 10, material('Si'), tables, what2)
    assert any(t == 'Mott' for r in res2 for t in r['scats']['type'])
    # command line, run from a directory next to tables/ like the reference ('../tables/mott/', extFunctions.py:44)
    import shutil
    shutil.copytree(os.path.join(os.path.dirname(__file__), 'data', 'tables', 'mott'), tmp_path / 'tables' / 'mott')
    os.makedirs(tmp_path / 'run', exist_ok=True)
    inp = tmp_path / 'input.file'
    inp.write_text('\n'.join([
        'num_el : 300', 'E0 : 20000 # eV', 'Emin : 15000', 'material : Si', 's_tilt : 70', 'thickness : 3000000',
        'elastic : Mott # Mott or Ruth_vanilla', 'mode : DS', 'Bethe_model : classical', 'diffMFP : no', 'gen_tables : yes',
        'tol_E : 1e-2', 'tol_W : 1e-3', 'maxScatt: 1000', 'Wc : 10',
        'electron_output : outcome, final_E', 'scatter_output : type, energy']))
    out = str(tmp_path / 'mott.h5')
    r = subprocess.run([sys.executable, '-m', 'mcdiscrete_b200', '-i', str(inp), '-o', out], cwd=str(tmp_path / 'run'),
                       env=dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', '')),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    back = read_store(out)
    assert len(back['els']) == 300
    assert 'Mott' in set(x.decode() if isinstance(x, bytes) else str(x) for x in back['scats']['type'].unique())
