"""ctypes binding of libmcd.so (include/mcd.h).  There is no fallback: if the CUDA library is
missing or no B200 is visible, every compute call raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MCD_LIB') or os.path.join(_HERE, 'libmcd.so')     # MCD_LIB: tuning experiments only

MAX_SHELLS = 8
MAX_CHANNELS = MAX_SHELLS + 3
MAX_TABLES = 1 + MAX_SHELLS
HIST_E_BINS, HIST_POLAR_BINS, HIST_AZ_BINS, HIST_DEPTH_BINS = 512, 180, 360, 512
HIST_TOTAL = HIST_E_BINS + HIST_POLAR_BINS + HIST_AZ_BINS + HIST_DEPTH_BINS

MODE_DS, MODE_CONT = 0, 1
BETHE = {'classical': 0, 'JL': 1, 'explicit': 2}
OUTCOMES = ('abs', 'bks', 'trsm', 'trdf', 'long')          # MCD_OUT_*
EL_BITS = {'outcome': 1, 'final_E': 2, 'final_dir': 4, 'last_pos': 8}
SC_BITS = {'type': 1, 'position': 2, 'pathl': 4, 'E': 8, 'E_loss': 16, 'pol_angle': 32, 'az_angle': 64,
           'energy': 128, 'direction': 256}
FLAG_HISTOGRAMS, FLAG_CHANNEL_COUNTS = 1, 2


class MCDError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('libmcd error %d: %s' % (code, msg))
        self.code = code


class Material(C.Structure):
    _fields_ = [('Z', C.c_int32), ('n_shells', C.c_int32), ('n_val', C.c_double), ('E_val', C.c_double),
                ('atnd', C.c_double), ('fermi_e', C.c_double), ('plasmon_e', C.c_double), ('bethe_k', C.c_double),
                ('shell_E', C.c_double * MAX_SHELLS), ('shell_n', C.c_double * MAX_SHELLS)]


class Params(C.Structure):
    _fields_ = [('E0', C.c_double), ('Emin', C.c_double), ('s_tilt_deg', C.c_double), ('thickness', C.c_double),
                ('Wc', C.c_double), ('tol_E', C.c_double), ('tol_W', C.c_double), ('maxScatt', C.c_int64),
                ('mode', C.c_int32), ('bethe', C.c_int32), ('elastic', C.c_int32), ('gryz_table_mode', C.c_int32),
                ('el_mask', C.c_uint32), ('scat_mask', C.c_uint32), ('flags', C.c_uint32), ('_pad', C.c_uint32),
                ('seed', C.c_uint64)]


class Tuning(C.Structure):
    _fields_ = [('handover_passes', C.c_int32), ('handover_threshold', C.c_int32), ('blocks_per_sm', C.c_int32),
                ('statistics_kernel', C.c_int32), ('tail_priority', C.c_int32)]


class ScatLayout(C.Structure):
    _fields_ = [('row_words', C.c_uint32), ('word_of', C.c_int32 * 16)]


class Counters(C.Structure):
    _fields_ = [('electrons_done', C.c_uint64), ('events', C.c_uint64), ('outcome', C.c_uint64 * 5),
                ('double_records', C.c_uint64), ('scat_rows_stored', C.c_uint64), ('scat_rows_dropped', C.c_uint64),
                ('nan_clamped', C.c_uint64), ('select_exact', C.c_uint64), ('channel', C.c_uint64 * MAX_CHANNELS)]


EL_ROW_DTYPE = np.dtype([('code', '<u4'), ('final_E', '<f4'), ('dx', '<f4'), ('dy', '<f4'), ('dz', '<f4'),
                         ('last_x', '<f4'), ('last_y', '<f4'), ('last_z', '<f4')])

# every symbol include/mcd.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int32)
_U64 = C.c_uint64
SYMBOLS = {
    'mcd_abi_version': (C.c_int, []),
    'mcd_last_error': (C.c_char_p, [_P]),
    'mcd_device_count': (C.c_int, [C.POINTER(C.c_int)]),
    'mcd_device_info': (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.POINTER(_U64)]),
    'mcd_create': (C.c_int, [C.POINTER(Params), C.POINTER(Material), C.c_int, C.POINTER(_P)]),
    'mcd_destroy': (C.c_int, [_P]),
    'mcd_clone': (C.c_int, [_P, C.POINTER(_P)]),
    'mcd_tuning_get': (C.c_int, [_P, C.POINTER(Tuning)]),
    'mcd_tuning_set': (C.c_int, [_P, C.POINTER(Tuning)]),
    'mcd_table_build': (C.c_int, [_P, C.c_int, _F, C.c_uint32, _F, C.c_uint32, C.c_double, C.c_double, _P]),
    'mcd_table_upload': (C.c_int, [_P, C.c_int, _F, C.c_uint32, _F, C.c_uint32, _F]),
    'mcd_mott_upload': (C.c_int, [_P, C.c_uint32, C.c_uint32] + [C.POINTER(C.c_double)] * 4),
    'mcd_table_shape': (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    'mcd_table_download': (C.c_int, [_P, C.c_int, _F]),
    'mcd_set_capacity': (C.c_int, [_P, _U64, _U64]),
    'mcd_scat_layout_get': (C.c_int, [_P, C.POINTER(ScatLayout)]),
    'mcd_reset_stats': (C.c_int, [_P, _P]),
    'mcd_run': (C.c_int, [_P, _U64, _U64, _P]),
    'mcd_sync': (C.c_int, [_P, _P]),
    'mcd_row_counts': (C.c_int, [_P, C.POINTER(_U64), C.POINTER(_U64)]),
    'mcd_drain_els': (C.c_int, [_P, _P, _U64, _U64, _P]),
    'mcd_drain_scats': (C.c_int, [_P, _P, _U64, _U64, _P]),
    'mcd_order_scats': (C.c_int, [_P, _P]),
    'mcd_drain_scats_ordered': (C.c_int, [_P, _P, _U64, _U64, _P]),
    'mcd_scat_offsets': (C.c_int, [_P, C.POINTER(_U64), _U64, _U64, _P]),
    'mcd_order_scats_columns': (C.c_int, [_P, _P]),
    'mcd_order_scats_columns_enqueue': (C.c_int, [_P, _P]),
    'mcd_scat_columns_layout': (C.c_int, [_P] + [C.POINTER(_U64)] * 5),
    'mcd_drain_scats_columns': (C.c_int, [_P, _P, _U64, _P]),
    'mcd_counters_get': (C.c_int, [_P, C.POINTER(Counters)]),
    'mcd_hist_get': (C.c_int, [_P, C.POINTER(C.c_int64)]),
    'mcd_stats_device_ptr': (C.c_int, [_P, C.POINTER(_P), C.POINTER(_U64)]),
    'mcd_stats_merge': (C.c_int, [_P, _P, _P]),
    'mcd_hist_ranges': (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    'mcd_last_kernel_ms': (C.c_int, [_P, C.POINTER(C.c_float)]),
    'mcd_launch_info': (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'mcd_last_launches': (C.c_int, [_P, C.POINTER(C.c_int)]),
    'mcd_measure_peaks': (C.c_int, [C.c_int, _F, _F, _F]),
    'mcd_test_sigmas': (C.c_int, [_P, _F, _U64, _F, _F]),
    'mcd_test_cum': (C.c_int, [_P, _F, _U64, _F, _I, _I, C.POINTER(C.c_double)]),
    'mcd_test_select': (C.c_int, [_P, _F, C.c_int, _F, _U64, _I, _I]),
    'mcd_test_ruth': (C.c_int, [_P, _F, _F, _U64, _F, _F]),
    'mcd_test_bincoll': (C.c_int, [_P, _F, _F, _U64, _F, _F]),
    'mcd_test_rotate': (C.c_int, [_P, _F, _F, _F, _F, _F, _U64, _F, _F]),
    'mcd_test_stopping': (C.c_int, [_P, C.c_int, _F, _U64, _F]),
    'mcd_test_lookup': (C.c_int, [_P, C.c_int, _F, _F, _U64, _F, _I, _I]),
    'mcd_test_mott': (C.c_int, [_P, _F, _F, _U64, _F, _F, _I, _I]),
    'mcd_test_step': (C.c_int, [_P, _F, _F, _U64, _F, _F]),
    'mcd_test_philox': (C.c_int, [_U64, _U64, C.c_uint32, C.POINTER(C.c_uint32), _F]),
}

_lib = None


def load():
    """Load libmcd.so and bind every symbol of include/mcd.h (raises if the library was not built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          '(nvcc, sm_100a).  mcdiscrete_b200 has no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if os.environ.get('MCD_LIB'):          # an older build variant in an A/B run: it simply lacks the call
                continue
            raise
        fn.restype = res
        fn.argtypes = args
    if lib.mcd_abi_version() != 1:
        raise ImportError('libmcd.so ABI version mismatch')
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc != 0:
        msg = load().mcd_last_error(ctx)
        raise MCDError(rc, (msg or b'').decode('utf-8', 'replace'))


def fptr(a):
    return a.ctypes.data_as(_F)


def iptr(a):
    return a.ctypes.data_as(_I)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)
