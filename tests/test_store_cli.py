"""CPU-only: the result store keeps the reference's keys / columns / order, and the CLI parses the reference's options."""
import os

import numpy as np
import pytest

from mcdiscrete_b200 import _lib
from mcdiscrete_b200.engine import RunResult, ScatRows
from mcdiscrete_b200.store import ResultStore, read_store, store_name, els_frame, scats_frame


def _fake_result(n_el=5):
    els = np.zeros(n_el, _lib.EL_ROW_DTYPE)
    els['code'] = [0, 1, 1 | 0x100, 4, 0]
    els['final_E'] = np.linspace(11000, 19000, n_el)
    els['dx'], els['dy'], els['dz'] = 0.1, 0.2, -0.9
    els['last_x'], els['last_y'], els['last_z'] = 1.0, 2.0, 3.0
    # layout for scatter_output = type, position, energy: words [rel_id, ev|ch<<24, x, y, z, energy]
    word_of = [1, 2, -1, -1, -1, -1, -1, 5, -1] + [-1] * 7
    rows = []
    for el, nev in enumerate([3, 1, 2, 2, 1]):
        for ev in range(nev):
            rows.append((el, ev | ((ev % 6) << 24), float(el), float(ev), 0.5, 20000. - 16.6 * ev))
    w = np.zeros((len(rows), 6), np.uint32)
    fl = w.view(np.float32)
    rng = np.random.default_rng(0).permutation(len(rows))      # append order is not trajectory order
    for k, r in enumerate(rows):
        i = rng[k]
        w[i, 0], w[i, 1] = r[0], r[1]
        fl[i, 2], fl[i, 3], fl[i, 4], fl[i, 5] = r[2], r[3], r[4], r[5]
    return RunResult(100, n_el, els, ScatRows(w, word_of, 100), {}, 0.0), rows


IP = dict(num_el=5, E0=20000., Emin=10000., material='Si', s_tilt=70., thickness=3e6, elastic='Ruth_vanilla', mode='DS',
          diffMFP=False, tol_E=1e-4, tol_W=1e-7, maxScatt=1000, Wc=10.,
          electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'], scatter_output=['type', 'position', 'energy'])
NAMES = ['Ruth', 'Moller', 'Gryz1s', 'Gryz2s', 'Gryz2p', 'Quinn']


def test_frames_follow_the_reference_column_contract():
    res, rows = _fake_result()
    e = els_frame(res, IP)
    assert list(e.columns) == ['outcome', 'final_E', 'dx', 'dy', 'dz', 'last_x', 'last_y', 'last_z']
    assert list(e['outcome']) == ['abs', 'bks', 'bks', 'long', 'abs']
    assert e['final_E'].dtype == np.float64
    e2 = els_frame(res, IP, double_record=True)                 # the reference's second 'abs' row
    assert list(e2['outcome']) == ['abs', 'bks', 'bks', 'abs', 'long', 'abs']
    s = scats_frame(res, IP, NAMES)
    assert list(s.columns) == ['type', 'energy', 'x', 'y', 'z']  # zipDict order (fileTools.py:72-91)
    assert len(s) == len(rows)
    assert list(s['x']) == [float(r[2]) for r in rows]           # regrouped per electron, event order
    assert list(s['y']) == [float(r[3]) for r in rows]
    assert list(s['type'][:3]) == ['Ruth', 'Moller', 'Gryz1s']


def test_store_roundtrip_parquet(tmp_path):
    res, rows = _fake_result()
    path = str(tmp_path / 'run.parquet')
    with ResultStore(path, IP, fmt='parquet') as st:
        st.append_result(res, NAMES)
        st.append_result(res, NAMES)                             # appendable
    back = read_store(path)
    assert back['input']['material'] == 'Si' and back['input']['num_el'] == '5'
    assert len(back['els']) == 10 and len(back['scats']) == 2 * len(rows)
    assert list(back['els'].columns) == ['outcome', 'final_E', 'dx', 'dy', 'dz', 'last_x', 'last_y', 'last_z']
    assert list(back['scats'].columns) == ['type', 'energy', 'x', 'y', 'z']


def test_store_hdf5_when_pytables_is_available(tmp_path):
    pytest.importorskip('tables')
    res, rows = _fake_result()
    path = str(tmp_path / 'run.h5')
    with ResultStore(path, IP, fmt='hdf5') as st:
        st.append_result(res, NAMES)
    back = read_store(path)
    assert len(back['els']) == 5 and list(back['scats'].columns) == ['type', 'energy', 'x', 'y', 'z']


def test_store_name_and_cli_options():
    from mcdiscrete_b200.doScatter import dealWithInput
    name = store_name(IP)
    assert name == ('../data/TRSM_diff:False_thick:3000000.0_mat:Si_mode:DS_elastic:Ruth_vanilla_tilt:70.0'
                    '_Emin:10000.0_E0:20000.0_tolE:0.0001_tolW:1e-07.h5')
    u, f, out, fmt, batch = dealWithInput(['-u', '-i', 'x.file', '--format', 'parquet', '--batch', '1000'])
    assert (u, f, out, fmt, batch) == (True, 'x.file', None, 'parquet', 1000)
    with pytest.raises(SystemExit):
        dealWithInput(['-h'])


def test_columnar_scats_view_equals_row_view(tmp_path):
    """ScatCols (the host view of mcd_drain_scats_columns: one contiguous array per column, no id / event words)
    exposes the same columns as ScatRows of the same rows in trajectory order, and the store writes the same file."""
    from mcdiscrete_b200.engine import ScatCols
    res, rows = _fake_result()
    sr = res.sorted_scats()                                  # row-major, trajectory order, with offsets
    n = len(sr)
    word_of = sr.word_of
    stride = ((4 * n + 255) // 256) * 256
    nf = 4                                                   # x, y, z, energy = row words 2..5
    block = np.zeros(nf * stride + n, np.uint8)
    for j in range(nf):
        block[j * stride:j * stride + 4 * n] = np.ascontiguousarray(sr.words[:, 2 + j]).view(np.uint8)
    block[nf * stride:] = (sr.words[:, 1] >> 24).astype(np.uint8)
    sc = ScatCols(block, n, stride, nf * stride, word_of, sr.first_id, sr.offsets)
    assert sc.keys() == sr.keys() and len(sc) == n
    for k in sr.keys():
        assert np.array_equal(sc[k], sr[k]), k
        assert sc[k].dtype == sr[k].dtype, k
    assert sc['energy'].flags['C_CONTIGUOUS'] and sc['type'].flags['C_CONTIGUOUS']
    res_c = RunResult(res.first_id, res.count, res.els, sc, {}, 0.0)
    assert res_c.sorted_scats() is sc
    a, b = str(tmp_path / 'rows.h5'), str(tmp_path / 'cols.h5')
    for path, r in ((a, res), (b, res_c)):
        with ResultStore(path, IP, fmt='hdf5') as st:
            st.append_result(r, NAMES)
    if open(a, 'rb').read(8) == b'\x89HDF\r\n\x1a\n' and not __import__('mcdiscrete_b200.store', fromlist=['x']).have_pytables():
        assert open(a, 'rb').read() == open(b, 'rb').read()  # the self-contained writer is deterministic
    ra, rb = read_store(a), read_store(b)
    assert ra['scats'].equals(rb['scats']) and ra['els'].equals(rb['els'])


@pytest.mark.parametrize('elastic', ['Ruth_vanilla_wDefl', 'Ruth_nigram_wDefl'])
def test_continuous_mode_elastic_labels_are_stored_in_full(tmp_path, elastic):
    """In continuous mode the scats `type` column holds the elastic label (scattering.py:802-840); the longest ones
    ('Ruth_vanilla_wDefl', 18 characters) must survive the fixed-width string column of the HDF5 store."""
    from mcdiscrete_b200.store import STR_WIDTH, scats_columns
    assert STR_WIDTH >= 18
    res, rows = _fake_result()
    ip = dict(IP, mode='cont', elastic=elastic, scatter_output=['type', 'energy'])
    # continuous rows: words [rel_id, ev|0<<24, energy]
    sr = res.sorted_scats()
    w = np.ascontiguousarray(sr.words[:, [0, 1, 5]])
    word_of = [1, -1, -1, -1, -1, -1, -1, 2, -1] + [-1] * 7
    res_c = RunResult(res.first_id, res.count, res.els, ScatRows(w, word_of, sr.first_id, sr.offsets), {}, 0.0)
    cols = scats_columns(res_c, ip, [elastic])
    assert set(cols['type'].tolist()) == {elastic.encode()}
    path = str(tmp_path / 'cont.h5')
    with ResultStore(path, ip, fmt='hdf5') as st:
        st.append_result(res_c, [elastic])
    back = read_store(path)
    assert set(back['scats']['type']) == {elastic}
    with pytest.raises(ValueError):
        scats_columns(res_c, dict(ip, elastic='x' * (STR_WIDTH + 1)), [elastic])
