"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/mcd.h
declares, fails loudly without a GPU, and the host-side mirror of the reference interface
(input parser, column contract, materials, table axis search) behaves like the reference."""
import ctypes as C
import json
import os
import sys
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    from mcdiscrete_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from mcdiscrete_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'mcd.h')).read()
    declared = set(re.findall(r'^\s*(?:int|const char\*)\s+(mcd_\w+)\s*\(', hdr, flags=re.M))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mcd_abi_version() == 1


def test_struct_sizes_match_header(lib):
    from mcdiscrete_b200 import _lib
    assert C.sizeof(_lib.Material) == 8 + 6 * 8 + 2 * 8 * 8
    assert C.sizeof(_lib.Params) == 7 * 8 + 8 + 4 * 4 + 4 * 4 + 8
    assert _lib.EL_ROW_DTYPE.itemsize == 32
    assert C.sizeof(_lib.Tuning) == 20
    assert C.sizeof(_lib.Counters) == 8 * (2 + 5 + 5 + 11)


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from mcdiscrete_b200 import _lib
    from mcdiscrete_b200.engine import Engine
    n = C.c_int()
    assert lib.mcd_device_count(C.byref(n)) == -3
    with pytest.raises(_lib.MCDError) as ei:
        Engine(dict(E0=2e4, Emin=1e4, material='Si', s_tilt=70., maxScatt=1000, mode='DS'))
    assert 'no CPU fallback' in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'mcdiscrete_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(import|from)\s+(mcd_oracle|oracle|ref_shim)', src, flags=re.M), f
                assert 'oracle/' not in src or f.endswith(('.cu', '.cuh')), f   # comments in kernels may cite it


def test_readInput_and_column_contract():
    from mcdiscrete_b200.fileTools import readInput, thingsToSave, zipDict, flat_columns
    ip = readInput(os.path.join(ROOT, 'tests', 'data', 'input.file'))
    assert ip['num_el'] == 10 and isinstance(ip['num_el'], int)
    assert ip['E0'] == 20000.0 and ip['material'] == 'Si' and ip['mode'] == 'DS'
    assert ip['gen_tables'] is True and ip['diffMFP'] is False
    assert ip['electron_output'] == ['outcome', 'final_E', 'final_dir', 'last_pos']
    assert ip['scatter_output'] == ['type', 'position', 'energy']
    assert ip['tol_W'] == 1e-7 and ip['maxScatt'] == 1000
    t = thingsToSave(ip['scatter_output'])
    t.addToList('position', [1., 2., 3.])
    t.addToList('type', 'Ruth')
    t.addToList('pathl', 5.)                      # not requested -> ignored (fileTools.py:219)
    assert t.dict == {'type': ['Ruth'], 'position': {'x': [1.], 'y': [2.], 'z': [3.]}, 'energy': []}
    merged = zipDict({}, t.dict)
    assert list(merged) == ['type', 'energy', 'x', 'y', 'z']          # flattened keys move to the end
    assert flat_columns(ip['scatter_output'], 'scats') == ['type', 'energy', 'x', 'y', 'z']
    assert flat_columns(ip['electron_output'], 'els') == ['outcome', 'final_E', 'dx', 'dy', 'dz',
                                                         'last_x', 'last_y', 'last_z']


def test_readInput_matches_the_reference_parser():
    """tests/golden/readInput.json = what the reference's own readInput (fileTools.py:8-46) returns for the files
    under tests/data/ (oracle/gen_golden_input.py): inline comments without a blank, extra float keys, yes/no flags,
    and its quirk of stripping blanks but not tabs (' \tCu' stays '\tCu')."""
    from mcdiscrete_b200.fileTools import readInput
    gold = json.load(open(os.path.join(GOLD, 'readInput.json')))
    assert set(gold) == {'input.file', 'input_variants.file'}
    for name, want in gold.items():
        got = readInput(os.path.join(ROOT, 'tests', 'data', name))
        assert got == want and list(got) == list(want), name          # same values, same key order
        for k in want:
            assert type(got[k]) is type(want[k]), (name, k)
    assert gold['input_variants.file']['material'] == '\tCu'


def test_thingsToSave_zipDict_and_column_order_match_the_reference():
    """tests/golden/columns.json = the reference's own thingsToSave / zipDict (fileTools.py:67-99, 193-242) and the
    column order of the DataFrame its listener stores (doScatter.py:112-117), for nine output selections driven
    through the same scripted addToList calls (oracle/gen_golden_columns.py)."""
    import copy
    from mcdiscrete_b200.fileTools import thingsToSave, zipDict, flat_columns
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import gen_golden_columns as gen              # only its call script (no reference import at test time)
    gold = json.load(open(os.path.join(GOLD, 'columns.json')))
    assert len(gold) == 9
    for key, c in gold.items():
        names = c['requested']
        dicts = gen.script(thingsToSave, names)
        assert dicts == c['per_electron'], key
        merged = {}
        for d in copy.deepcopy(dicts):
            merged = zipDict(merged, d)
        assert merged == c['zipped'] and list(merged) == list(c['zipped']), key
        which = 'els' if set(names) & {'outcome', 'final_E', 'final_dir', 'last_pos'} else 'scats'
        assert flat_columns(names, which) == c['frame_columns'], key


def test_material_matches_reference():
    from mcdiscrete_b200.material import material
    F = json.load(open(os.path.join(GOLD, 'formulas.json')))
    for name, g in F.items():
        m = material(name)
        assert m.atnd == pytest.approx(g['atnd'], rel=1e-14)
        assert m.fermi_e == pytest.approx(g['fermi_e'], rel=1e-14)
        assert m.plasmon_e == pytest.approx(g['plasmon_e'], rel=1e-14)
        assert m.channel_names[0] == 'Ruth' and m.channel_names[-1] == 'Quinn'


@pytest.mark.parametrize('key', ['Si_15k_small', 'Al_500_small', 'Cu_10k_small', 'Si_10k_mid', 'Al_500_mid'])
def test_host_axis_search_matches_reference(key):
    """findWseries / findEseries (host side of the table build) against the reference's grids."""
    from mcdiscrete_b200.material import material
    from mcdiscrete_b200.probTables import probTable
    ax = json.load(open(os.path.join(GOLD, 'table_axes.json')))[key]
    ip = ax['input']
    mat = material(ip['material'])
    T = probTable('Moller', mat.params['name_val'], (ip['Emin'], ip['E0']), ip['tol_E'], ip['tol_W'], mat, Wc=ip['Wc'])
    T.set_Ws(); T.set_Es()
    assert [len(T.Es), len(T.Ws)] == ax['moller']
    G = probTable('Gryzinski', mat.params['name_s'][-1], (ip['Emin'], ip['E0']), ip['tol_E'], ip['tol_W'], mat)
    G.set_Ws(); G.set_Es()
    assert [G.shell, len(G.Es), len(G.Ws)] == ax['gryz'][0]
    if key.endswith('small'):
        g = np.load(os.path.join(GOLD, 'tables_%s.npz' % key))
        assert np.array_equal(T.Ws, g['moller_Ws']) and np.array_equal(T.Es, g['moller_Es'])
        assert np.array_equal(G.Ws, g['gryz0_Ws']) and np.array_equal(G.Es, g['gryz0_Es'])


@pytest.mark.parametrize('key', ['C2_Si', 'C3_Cu', 'C4_Au'])
def test_host_axis_search_default_tolerance(key):
    """The host axis search at BASELINE's default tolerances (tol_E 1e-4, tol_W 1e-7) against the reference's grids
    (SHA-1 of the float32 axes, oracle/gen_golden_defaults.py)."""
    import hashlib
    from mcdiscrete_b200.material import material
    from mcdiscrete_b200.probTables import probTable
    ax = json.load(open(os.path.join(GOLD, 'table_axes_default.json')))[key]
    ip = ax['input']
    mat = material(ip['material'])
    sha = lambda a: hashlib.sha1(np.ascontiguousarray(a, np.float32).tobytes()).hexdigest()
    for nm, T in (('moller', probTable('Moller', mat.params['name_val'], (ip['Emin'], ip['E0']), ip['tol_E'], ip['tol_W'],
                                       mat, Wc=ip['Wc'])),
                  ('gryz0', probTable('Gryzinski', mat.params['name_s'][-1], (ip['Emin'], ip['E0']), ip['tol_E'],
                                      ip['tol_W'], mat))):
        T.set_Ws(); T.set_Es()
        a = ax[nm]
        assert [T.shell, len(T.Es), len(T.Ws)] == [a['shell'], a['nE'], a['nW']]
        assert T.Es.dtype == np.float32 and T.Ws.dtype == np.float32
        assert sha(T.Es) == a['Es_sha1'] and sha(T.Ws) == a['Ws_sha1']
        assert not T.exhausted


def test_argument_checks_precede_the_device(lib):
    """mcd_create rejects what the packed row format cannot hold before it touches a device: the event index of a
    scats row / parked electron is a 24-bit field, so maxScatt >= 2^24 is MCD_ERR_INVALID (not silent corruption)."""
    from mcdiscrete_b200 import _lib
    from mcdiscrete_b200.engine import Engine
    ip = dict(E0=2e4, Emin=1e4, material='Si', s_tilt=70., mode='DS')
    for bad in (1 << 24, 10 ** 9):
        with pytest.raises(_lib.MCDError) as ei:
            Engine(dict(ip, maxScatt=bad))
        assert ei.value.code == -1 and '24-bit' in str(ei.value)


def test_multiScatter_engine_key_covers_every_parameter():
    """Two worker calls that differ in ANY field feeding mcd_params / mcd_material must not share a context."""
    from mcdiscrete_b200.material import material
    from mcdiscrete_b200.multiScatter import _run_key
    tts = {'el_output': ['outcome'], 'scat_output': []}
    ip = dict(E0=2e4, Emin=1e4, material='Si', s_tilt=70., maxScatt=1000, mode='DS', elastic='Ruth_vanilla',
              thickness=3e6, Wc=10., tol_E=1e-2, tol_W=1e-3, seed=1, Bethe_model='classical')
    mat = material('Si')
    base = _run_key(ip, mat, tts)
    for k, v in dict(elastic='Ruth_nigram', thickness=2000., seed=2, tol_E=1e-3, tol_W=1e-4, Wc=20., maxScatt=10,
                     s_tilt=0., Emin=5e3, E0=3e4, mode='cont').items():
        assert _run_key(dict(ip, **{k: v}), mat, tts) != base, k
    assert _run_key(ip, material('Cu'), tts) != base                       # the material OBJECT, not inputPar['material']
    m2 = material('Si'); m2.params['Es']['2p'] = 101.
    assert _run_key(ip, m2, tts) != base
    assert _run_key(ip, mat, {'el_output': ['outcome'], 'scat_output': ['type']}) != base
    c = dict(ip, mode='cont')
    assert _run_key(dict(c, Bethe_model='JL'), mat, tts) != _run_key(c, mat, tts)
    # continuous mode: the _wDefl labels run the plain forms (scattering.py:781-789) -> same physics, same key
    assert _run_key(dict(c, elastic='Ruth_vanilla_wDefl'), mat, tts) == _run_key(c, mat, tts)


def test_mottTable_host_mirror_matches_the_reference(tmp_path):
    """mcdiscrete_b200.extFunctions.mottTable / pickMottTable against the arrays and picks the reference's own
    classes produced on the same (synthetic, ioffe-layout) file -- tests/golden/mott_Si.npz by
    oracle/gen_golden_mott.py (extFunctions.py:21-153, scattering.py:96-122)."""
    import random
    import shutil
    from mcdiscrete_b200.extFunctions import mottTable, pickMottTable
    from mcdiscrete_b200.engine import elastic_kind, ELASTIC_MOTT
    g = np.load(os.path.join(GOLD, 'mott_Si.npz'))
    work = tmp_path / 'mott'
    shutil.copytree(os.path.join(ROOT, 'tests', 'data', 'tables', 'mott'), work)
    mt = mottTable(14, 'Si', str(work) + os.sep)
    assert (mt.sizeE, mt.sizeTheta) == g['probTable'].shape
    mt.readIoffe().toProbTable()
    for k in ('Es', 'sigmas', 'thetas', 'dCS', 'probTable'):
        assert np.array_equal(np.asarray(getattr(mt, k), float), g[k]), k
    saved = random.random
    try:
        for E, u, c2 in zip(g['pick_E'], g['pick_u'], g['pick_c2']):
            random.random = lambda u=u: float(u)
            assert pickMottTable(mt, float(E)) == c2
    finally:
        random.random = saved
    # the float32 memmap round trip (mapToMemory / readFromMemory, :112-145)
    mt.mapToMemory()
    m2 = mottTable(14, 'Si', str(work) + os.sep).readFromMemory()
    assert m2.probTable.dtype == np.float32 and np.array_equal(m2.probTable, g['probTable'].astype(np.float32))
    assert np.array_equal(m2.Es, g['Es'])
    with pytest.raises(ValueError):
        mottTable(13, 'Si', str(work) + os.sep)                  # Z check (:58-60)
    with pytest.raises(FileNotFoundError):
        mottTable(29, 'Cu', str(work) + os.sep)
    assert elastic_kind('Mott', True) == ELASTIC_MOTT
    with pytest.raises(NotImplementedError):
        elastic_kind('Mott', False)                              # AttributeError in the reference (golden records it)
    assert 'tableMott' in str(g['cont_behaviour'])
