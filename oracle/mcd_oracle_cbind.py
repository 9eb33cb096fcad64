"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_build/libmcd_oracle.so (oracle/mcd_oracle_c.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

import mcd_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, '_build', 'libmcd_oracle.so')


class CMat(C.Structure):
    _fields_ = [('Z', C.c_int), ('nshell', C.c_int), ('n_val', C.c_double), ('E_val', C.c_double), ('atnd', C.c_double),
                ('Ef', C.c_double), ('Epl', C.c_double), ('bethe_k', C.c_double), ('bohr_r', C.c_double),
                ('Es', C.c_double * 8), ('ns', C.c_double * 8)]


class CTab(C.Structure):
    _fields_ = [('cdf', C.POINTER(C.c_float)), ('Es', C.POINTER(C.c_float)), ('Ws', C.POINTER(C.c_float)),
                ('nE', C.c_int), ('nW', C.c_int)]


class CPar(C.Structure):
    _fields_ = [('E0', C.c_double), ('Emin', C.c_double), ('tilt_deg', C.c_double), ('thickness', C.c_double),
                ('Wc', C.c_double), ('maxScatt', C.c_long), ('mode', C.c_int), ('bethe', C.c_int),
                ('seed', C.c_ulonglong), ('elastic', C.c_int), ('mott_nE', C.c_int), ('mott_nT', C.c_int),
                ('mott_Es', C.POINTER(C.c_double)), ('mott_sig', C.POINTER(C.c_double)),
                ('mott_cdf', C.POINTER(C.c_double)), ('mott_theta', C.POINTER(C.c_double))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, 'mcd_oracle_c.c')):
            subprocess.check_call(['make', '-C', HERE])
        _lib = C.CDLL(SO)
        _lib.orc_run.restype = C.c_int
    return _lib


def _tab(T):
    cdf = np.ascontiguousarray(T.table, np.float32)
    Es = np.ascontiguousarray(T.Es, np.float32)
    Ws = np.ascontiguousarray(T.Ws, np.float32)
    t = CTab(cdf.ctypes.data_as(C.POINTER(C.c_float)), Es.ctypes.data_as(C.POINTER(C.c_float)),
             Ws.ctypes.data_as(C.POINTER(C.c_float)), len(Es), len(Ws))
    t._keep = (cdf, Es, Ws)
    return t


def run(inputPar, first_id, count, seed, tables=None, mat=None, nthreads=1, event_rows=0):
    """Returns dict(outcome, n_rows, final_E, final_dir, last_pos, n_ev[, events])."""
    mat = mat or o.Material(inputPar['material'])
    m = CMat(mat.Z, len(mat.Es), mat.n_val, mat.E_val, mat.atnd, mat.fermi_e, mat.plasmon_e, mat.bethe_k, o.BOHR_R)
    for i, (e, n) in enumerate(zip(mat.Es, mat.ns)):
        m.Es[i], m.ns[i] = e, n
    ds = inputPar['mode'] == 'DS'
    el = inputPar.get('elastic', 'Ruth_vanilla')
    mott = el == 'Mott'
    if mott and not ds:
        raise ValueError("elastic = 'Mott' exists in DS mode only")
    if not ds:                                        # scattering.py:781-789: the _wDefl branches are unreachable
        el = 'Ruth_vanilla' if 'vanilla' in el else 'Ruth_nigram'
    p = CPar(inputPar['E0'], inputPar['Emin'], inputPar['s_tilt'], inputPar['thickness'], inputPar.get('Wc', 10.),
             int(inputPar['maxScatt']), 0 if ds else 1,
             {'classical': 0, 'JL': 1, 'explicit': 2}[inputPar.get('Bethe_model', 'classical')], seed,
             4 if mott else o.ELASTIC_MODELS.index(el))
    if mott:
        M = tables['Mott']
        keep = [np.ascontiguousarray(a, np.float64) for a in (M.Es, M.sigmas, M.probTable, M.thetas)]
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        p.mott_nE, p.mott_nT = len(keep[0]), len(keep[3])
        p.mott_Es, p.mott_sig, p.mott_cdf, p.mott_theta = dp(keep[0]), dp(keep[1]), dp(keep[2]), dp(keep[3])
        p._keep = keep
    tm = _tab(tables['Moller']) if ds else None
    tg = _tab(tables['Gryz'][0]) if ds else None
    els = np.empty((count, 9), np.float64)
    nev = np.empty(count, np.int32)
    ev = np.empty((event_rows, 8), np.float64) if event_rows else None
    rc = lib().orc_run(C.byref(p), C.byref(m), C.byref(tm) if ds else None, C.byref(tg) if ds else None,
                       C.c_uint64(first_id), C.c_uint64(count), els.ctypes.data_as(C.c_void_p),
                       nev.ctypes.data_as(C.c_void_p), C.c_int(nthreads),
                       ev.ctypes.data_as(C.c_void_p) if event_rows else None, C.c_long(event_rows))
    if rc:
        raise RuntimeError('orc_run failed: %d' % rc)
    out = dict(outcome=els[:, 0].astype(np.int8), n_rows=els[:, 1].astype(np.int8), final_E=els[:, 2],
               final_dir=els[:, 3:6], last_pos=els[:, 6:9], n_ev=nev)
    if event_rows:
        out['events'] = ev[:min(event_rows, int(nev.sum()))]
    return out
