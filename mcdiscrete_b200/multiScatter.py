"""Drop-in workers with the reference's names and signatures (MC/tools/multiScatter.py).

    multiTraj_DS(inputPar, numTraj, material, tables, thingsToSave)       multiScatter.py:66-116
    multiTraj_cont(inputPar, numTraj, material, thingsToSave, tables)     multiScatter.py:171-229
    scatterMultiEl_DS / scatterMultiEl_cont                               multiScatter.py:15-61, 118-168

Each returns what the reference worker returns -- a list (len numTraj) of
``{'els': {...}, 'scats': {...}}`` dict-of-lists, vector columns nested exactly as
``fileTools.thingsToSave`` nests them -- but the trajectories run on the GPU in one batch.
For bulk runs use :class:`mcdiscrete_b200.engine.Engine` (columnar arrays, no per-electron dicts).

Deviations from the reference, all documented in DESIGN.md:
  * RNG: Philox4x32-10 keyed on (seed, electron id, event) instead of ``random.seed(pid*time)`` per job
    (multiScatter.py:94 restarts identical streams within one second).  Successive calls
    continue with fresh electron ids; ``inputPar['seed']`` (default 12345) selects the stream.
  * a 'bks'/'trsm' electron whose exit event also crosses Emin is reported once (the reference appends a
    second 'abs' row) unless ``inputPar['reference_double_record']`` is true.
  * 'long' electrons carry exactly maxScatt events (the reference leaves one dangling position, Q6).
  * multiTraj_cont works for all three Bethe models (the reference's crashes).
"""
import numpy as np

from . import _lib
from .engine import Engine
from .material import material as Material

_engines = {}


def _run_key(inputPar, material, thingsToSave):
    """Every field that feeds mcd_params / mcd_material (Engine.__init__): two calls with different keys must never
    share a context."""
    from .engine import elastic_kind
    ds = 'DS' in inputPar['mode']
    prm = material.params
    mat_key = (getattr(material, 'species', inputPar['material']), prm['Z'], prm['n_val'], prm['E_val'],
               tuple(prm['name_s']), tuple(float(prm['Es'][s]) for s in prm['name_s']),
               tuple(float(prm['ns'][s]) for s in prm['name_s']), float(prm.get('bethe_k', 0.0)),
               float(material.atnd), float(material.fermi_e), float(material.plasmon_e))
    return (mat_key, 'DS' if ds else 'cont', inputPar.get('Bethe_model', 'classical') if not ds else None,
            elastic_kind(inputPar.get('elastic', 'Ruth_vanilla'), ds),
            float(inputPar['E0']), float(inputPar['Emin']), float(inputPar['s_tilt']),
            float(inputPar.get('thickness', 3e6)), int(inputPar['maxScatt']), float(inputPar.get('Wc', 10.)),
            float(inputPar.get('tol_E', 0)), float(inputPar.get('tol_W', 0)),
            tuple(thingsToSave['el_output']), tuple(thingsToSave['scat_output']), inputPar.get('seed'))


def _engine_for(inputPar, material, thingsToSave, tables=None):
    key = _run_key(inputPar, material, thingsToSave)
    if tables:
        for T in [tables.get('Moller')] + list(tables.get('Gryz', [])):
            eng = getattr(T, 'engine', None)
            if eng is not None and _compatible(eng, key):
                return eng
    eng = _engines.get(key)
    if eng is None:
        ip = dict(inputPar, electron_output=list(thingsToSave['el_output']),
                  scatter_output=list(thingsToSave['scat_output']))
        eng = Engine(ip, material=material)
        if 'DS' in inputPar['mode']:
            if tables and getattr(tables.get('Moller'), 'engine', None) is not None:
                # tables built for another context: move the rows over (device -> host -> device)
                for slot, T in enumerate([tables['Moller']] + list(tables['Gryz'])):
                    eng.table_upload(slot, T.Es, T.Ws, T.engine.table_download(T.slot))
                eng.tables = tables
            elif tables:
                # the reference's own probTable objects (masked memmap rows, probTables.py:513-535)
                for slot, T in enumerate([tables['Moller'], tables['Gryz'][0]]):
                    eng.table_upload(slot, T.Es, T.Ws, np.ma.filled(T.table, 0))
                eng.tables = tables
            if tables and tables.get('Mott') is not None and inputPar.get('elastic') == 'Mott':
                M = tables['Mott']              # ours or the reference's mottTable (extFunctions.py:21-153)
                eng.mott_upload(M.Es, M.sigmas, np.asarray(M.probTable, dtype=np.float64), M.thetas)
        _engines[key] = eng
    return eng


def _compatible(eng, key):
    """The engine the tables were built on may be reused only if it was created from the same parameters (elastic
    model, thickness, Bethe model, seed and tolerances included) and the same material."""
    try:
        return eng.gryz_mode == 'outermost' and \
            _run_key(eng.inputPar, eng.material, {'el_output': eng.el_cols, 'scat_output': eng.scat_cols}) == key
    except (KeyError, NotImplementedError):
        return False


def _per_electron_dicts(eng, res, inputPar, thingsToSave):
    """Columnar RunResult -> the reference's list of nested dict-of-lists (multiScatter.py:112)."""
    el_cols, sc_cols = list(thingsToSave['el_output']), list(thingsToSave['scat_output'])
    names = eng.channel_names
    cont = 'DS' not in inputPar['mode']
    sc = res.sorted_scats()
    n = res.count
    have_sc = len(sc) > 0
    if have_sc:
        bounds = sc.offsets.astype(np.int64)
    double = bool(inputPar.get('reference_double_record', False))
    out = []
    for i in range(n):
        r = res.els[i]
        outcomes = [_lib.OUTCOMES[int(r['code']) & 0xff]]
        if double and (int(r['code']) & 0x100):
            outcomes.append('abs')
        els = {}
        for c in el_cols:
            k = len(outcomes)
            if c == 'outcome':
                els[c] = list(outcomes)
            elif c == 'final_E':
                els[c] = [float(r['final_E'])] * k
            elif c == 'final_dir':
                els[c] = {'dx': [float(r['dx'])] * k, 'dy': [float(r['dy'])] * k, 'dz': [float(r['dz'])] * k}
            elif c == 'last_pos':
                els[c] = {'last_x': [float(r['last_x'])] * k, 'last_y': [float(r['last_y'])] * k,
                          'last_z': [float(r['last_z'])] * k}
        scd = {}
        lo, hi = (bounds[i], bounds[i + 1]) if have_sc else (0, 0)
        for c in sc_cols:
            if cont and c in ('position', 'E'):
                scd[c] = {'x': [], 'y': [], 'z': []} if c == 'position' else []
            elif c == 'type':
                scd[c] = [inputPar.get('elastic', 'Ruth_vanilla')] * (hi - lo) if cont else \
                    [names[t] for t in sc['type'][lo:hi]]
            elif c == 'position':
                scd[c] = {k: sc[k][lo:hi].astype(float).tolist() for k in ('x', 'y', 'z')}
            elif c == 'direction':
                scd[c] = [np.asarray(v, float) for v in sc['direction'][lo:hi]]
            else:
                scd[c] = sc[c][lo:hi].astype(float).tolist()
        out.append({'els': els, 'scats': scd})
    return out


def multiTraj_DS(inputPar, numTraj, material, tables, thingsToSave):
    """multiScatter.py:66-116."""
    eng = _engine_for(inputPar, material, thingsToSave, tables)
    res = eng.run(count=numTraj, ordered=True)
    return _per_electron_dicts(eng, res, inputPar, thingsToSave)


def multiTraj_cont(inputPar, numTraj, material, thingsToSave, tables):
    """multiScatter.py:171-229 (argument order as in the reference)."""
    if inputPar.get('Bethe_model', 'classical') not in _lib.BETHE:
        raise SystemExit('! I did not understand your choice of Bethe model in multiScatter')
    eng = _engine_for(inputPar, material, thingsToSave, None)
    res = eng.run(count=numTraj, ordered=True)
    return _per_electron_dicts(eng, res, inputPar, thingsToSave)


def scatterMultiEl_DS(inputPar, tables, listParToSave, outputQ, count):
    """multiScatter.py:15-61: runs inputPar['num_el'] electrons; the reference overwrites outputQ['els'] /
    outputQ['scats'] per electron (:54-55), leaving the LAST electron's dicts -- same here."""
    res = multiTraj_DS(inputPar, inputPar['num_el'], Material(inputPar['material']), tables, listParToSave)
    outputQ['els'] = res[-1]['els']
    outputQ['scats'] = res[-1]['scats']


def scatterMultiEl_cont(inputPar, thingsToSave, output, num, count):
    """multiScatter.py:118-168: `num` electrons; pickled dicts are put on the two queues."""
    import pickle
    from .fileTools import zipDict
    res = multiTraj_cont(inputPar, num, Material(inputPar['material']), thingsToSave, {})
    els, scats = {}, {}
    for r in res:
        els = zipDict(els, r['els'])
        scats = zipDict(scats, r['scats'])
    output['electrons'].put(pickle.dumps(els, protocol=2))
    output['scatterings'].put(pickle.dumps(scats, protocol=2))
