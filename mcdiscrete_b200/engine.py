"""Engine: one libmcd context on one GPU -- the batched replacement for the reference's worker pool
(MC/doScatter.py:22-95 MapScatterer + MC/tools/multiScatter.py:66-116,171-229 workers).

PyTorch is used for plumbing only (pinned host buffers, the current CUDA stream, torch.distributed);
all physics runs inside libmcd.so."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, fptr, iptr, f32
from .material import material as Material, as_c_material

DEFAULT_SEED = 12345


def _mask(names, bits):
    m = 0
    for n in names:
        if n not in bits:
            raise ValueError('unknown output column %r (known: %s)' % (n, ', '.join(bits)))
        m |= bits[n]
    return m


class ScatRows:
    """The drained scats rows: a (n_rows, row_words) uint32 array plus lazily evaluated columns.

    Column access is dict-like (``sc['energy']``, ``'x' in sc``, ``sc.keys()``).  Float columns are strided VIEWS
    into the packed rows (no copy); 'el' (absolute electron id), 'ev' (event index) and 'type' (channel index in
    sigma insertion order) are decoded on first use."""

    _FLOAT = {1: 'position', 2: 'pathl', 3: 'E', 4: 'E_loss', 5: 'pol_angle', 6: 'az_angle', 7: 'energy',
              8: 'direction'}

    def __init__(self, words, word_of, first_id, offsets=None):
        self.words, self.word_of, self.first_id = words, word_of, int(first_id)
        self.offsets = offsets        # set when the rows are in trajectory order: electron i owns rows offsets[i]:offsets[i+1]
        self._cache = {}
        self._names = ['el', 'ev']
        if word_of[0] >= 0:
            self._names.append('type')
        for b, nm in self._FLOAT.items():
            if word_of[b] >= 0:
                self._names += ['x', 'y', 'z'] if nm == 'position' else [nm]

    def __len__(self):
        return len(self.words)

    def __bool__(self):
        return True

    def keys(self):
        return list(self._names)

    def __contains__(self, k):
        return k in self._names

    def items(self):
        return [(k, self[k]) for k in self._names]

    def __getitem__(self, k):
        if k in self._cache:
            return self._cache[k]
        if k not in self._names:
            raise KeyError(k)
        w = self.words
        fl = w.view(np.float32)
        if k == 'el':
            v = w[:, 0].astype(np.int64) + self.first_id
        elif k == 'ev':
            v = (w[:, 1] & 0xffffff).astype(np.int32)
        elif k == 'type':
            v = (w[:, 1] >> 24).astype(np.int8)
        elif k in ('x', 'y', 'z'):
            v = fl[:, self.word_of[1] + 'xyz'.index(k)]
        elif k == 'direction':
            v = fl[:, self.word_of[8]:self.word_of[8] + 3]
        else:
            b = [b for b, nm in self._FLOAT.items() if nm == k][0]
            v = fl[:, self.word_of[b]]
        self._cache[k] = v
        return v


class ScatCols:
    """scats rows drained COLUMN by column in trajectory order (mcd_order_scats_columns): same dict-like interface
    as ScatRows, but every float column is a contiguous float32 array and 'type' a contiguous int8 array -- what
    the reference finally stores (DataFrame columns, doScatter.py:112-117).  'el' and 'ev' are not transferred;
    they are rebuilt from `offsets` (electron i owns rows offsets[i]:offsets[i+1]) on first use."""

    def __init__(self, block, n_rows, col_stride, type_offset, word_of, first_id, offsets):
        self.block, self.n, self.word_of, self.first_id, self.offsets = block, int(n_rows), word_of, int(first_id), offsets
        self._stride, self._type_off = int(col_stride), int(type_offset)
        self._cache = {}
        self._names = ['el', 'ev']
        if word_of[0] >= 0:
            self._names.append('type')
        for b, nm in ScatRows._FLOAT.items():
            if word_of[b] >= 0:
                self._names += ['x', 'y', 'z'] if nm == 'position' else [nm]

    def __len__(self):
        return self.n

    def __bool__(self):
        return True

    def keys(self):
        return list(self._names)

    def __contains__(self, k):
        return k in self._names

    def items(self):
        return [(k, self[k]) for k in self._names]

    def _word(self, w):
        a = (w - 2) * self._stride
        return self.block[a:a + 4 * self.n].view(np.float32)

    def __getitem__(self, k):
        if k in self._cache:
            return self._cache[k]
        if k not in self._names:
            raise KeyError(k)
        if k in ('el', 'ev'):
            off = np.asarray(self.offsets, np.int64)
            nev = np.diff(off)
            el = np.repeat(np.arange(len(nev), dtype=np.int64), nev)
            v = el + self.first_id if k == 'el' else (np.arange(self.n, dtype=np.int64) - off[:-1][el]).astype(np.int32)
        elif k == 'type':
            v = self.block[self._type_off:self._type_off + self.n].view(np.int8)
        elif k in ('x', 'y', 'z'):
            v = self._word(self.word_of[1] + 'xyz'.index(k))
        elif k == 'direction':
            v = np.stack([self._word(self.word_of[8] + i) for i in range(3)], 1)
        else:
            b = [b for b, nm in ScatRows._FLOAT.items() if nm == k][0]
            v = self._word(self.word_of[b])
        self._cache[k] = v
        return v


class RunResult:
    """Columnar output of one run.

    els   : structured array (code, final_E, dx, dy, dz, last_x, last_y, last_z), row i = electron first_id+i
    scats : ScatRows (dict-like columns in append order, with 'el' = absolute electron id and 'ev' = event
            index) or {} when no scatter_output was requested
    """

    def __init__(self, first_id, count, els, scats, counters, kernel_ms):
        self.first_id, self.count = first_id, count
        self.els, self.scats, self.counters, self.kernel_ms = els, scats, counters, kernel_ms

    @property
    def outcome(self):
        return (self.els['code'] & 0xff).astype(np.int8)

    @property
    def double_record(self):
        return (self.els['code'] & 0x100) != 0

    def sorted_scats(self):
        """scats columns regrouped per electron in event order (the reference's row order)."""
        if not len(self.scats) or self.scats.offsets is not None:
            return self.scats
        w = self.scats.words
        order = np.lexsort((w[:, 1] & 0xffffff, w[:, 0]))
        off = np.concatenate([[0], np.cumsum(np.bincount(w[:, 0].astype(np.int64), minlength=self.count))]).astype(np.uint64)
        return ScatRows(w[order], self.scats.word_of, self.scats.first_id, off)


ELASTIC_MODELS = ('Ruth_vanilla', 'Ruth_vanilla_wDefl', 'Ruth_nigram', 'Ruth_nigram_wDefl')
ELASTIC_MOTT = 4          # MCD_ELASTIC_MOTT


def elastic_kind(elastic, ds):
    """input.file `elastic` label -> the cross-section the reference evaluates for it (MCD_ELASTIC_* of mcd.h).
    DS (scattering.py:483-491): Ruth_vanilla -> ruther_sigma, *vanilla_wDefl -> ruther_sigma_wDefl (raises NameError
    in the reference because scattering.py:14 does not import it; the intended function is used here), Ruth_nigram ->
    ruther_N_sigma, *nigram_wDefl -> ruther_N_sigma_wDefl.  Continuous mode (scattering.py:781-789) tests the
    substrings 'vanilla', 'vanilla_wDefl', 'nigram', 'nigram_wDefl' in that order, so its _wDefl branches are
    unreachable and those labels run the plain forms (tests/golden/elastic_formulas.json records both facts).
    'Mott' (DS only, scattering.py:466-467, :493, :588-591) takes the elastic cross-section and the polar angle from
    the table genTables reads (probTables.py:601-613); the reference's continuous classes never receive that table
    (scattering.py:794 raises AttributeError, recorded in tests/golden/mott_Si.npz), so it is refused there.
    Browning does not exist in the reference's code."""
    if elastic == 'Mott':
        if not ds:
            raise NotImplementedError("elastic = 'Mott' with a continuous mode: the reference raises AttributeError "
                                      "(scattering.py:794: the class never stores tableMott)")
        return ELASTIC_MOTT
    if elastic not in ELASTIC_MODELS:
        raise NotImplementedError("elastic = %r: only %s and Mott are on the path"
                                  % (elastic, ', '.join(ELASTIC_MODELS)))
    if not ds:
        elastic = 'Ruth_vanilla' if 'vanilla' in elastic else 'Ruth_nigram'
    return ELASTIC_MODELS.index(elastic)


class Engine:
    def __init__(self, inputPar, material=None, device=None, seed=None, gryz_table='outermost',
                 histograms=True, channel_counts=False):
        self.lib = _lib.load()
        self.inputPar = dict(inputPar)
        self.material = material if material is not None else Material(inputPar['material'])
        self.gryz_mode = gryz_table
        if device is None:
            device = 0
            try:
                import torch
                if torch.cuda.is_available():
                    device = torch.cuda.current_device()
            except Exception:
                pass
        self.device = int(device)
        ip = self.inputPar
        if ip.get('diffMFP', False):
            raise NotImplementedError('diffMFP: yes crashes in the reference (scattering.py:512) and is not implemented')
        p = _lib.Params()
        p.E0, p.Emin = float(ip['E0']), float(ip['Emin'])
        p.s_tilt_deg = float(ip['s_tilt'])
        p.thickness = float(ip.get('thickness', 3e6))
        p.Wc = float(ip.get('Wc', 10.))
        p.tol_E, p.tol_W = float(ip.get('tol_E', 0)), float(ip.get('tol_W', 0))
        p.maxScatt = int(ip['maxScatt'])
        mode = ip['mode']
        p.mode = _lib.MODE_DS if 'DS' in mode else _lib.MODE_CONT
        if p.mode == _lib.MODE_CONT and 'cont' not in mode:
            raise ValueError("mode must be 'DS' or 'cont' (got %r)" % mode)
        bm = ip.get('Bethe_model', 'classical')
        if bm not in _lib.BETHE:
            raise ValueError('! I did not understand your choice of Bethe model: %r' % bm)
        p.bethe = _lib.BETHE[bm]
        p.elastic = elastic_kind(ip.get('elastic', 'Ruth_vanilla'), p.mode == _lib.MODE_DS)
        p.gryz_table_mode = 1 if gryz_table == 'per_shell' else 0
        self.el_cols = list(ip.get('electron_output', ['outcome', 'final_E', 'final_dir', 'last_pos']))
        self.scat_cols = list(ip.get('scatter_output', []))
        p.el_mask = _mask(self.el_cols, _lib.EL_BITS)
        scat_cols = self.scat_cols
        if p.mode == _lib.MODE_CONT:      # continuous mode never records position / E (scattering.py:802-840)
            scat_cols = [c for c in scat_cols if c not in ('position', 'E')]
        p.scat_mask = _mask(scat_cols, _lib.SC_BITS)
        p.flags = (_lib.FLAG_HISTOGRAMS if histograms else 0) | (_lib.FLAG_CHANNEL_COUNTS if channel_counts else 0)
        p.seed = int(seed if seed is not None else ip.get('seed', DEFAULT_SEED)) & 0xffffffffffffffff
        self.params = p
        self.cmat = as_c_material(self.material)
        self.ctx = C.c_void_p()
        check(self.lib.mcd_create(C.byref(p), C.byref(self.cmat), self.device, C.byref(self.ctx)))
        lay = _lib.ScatLayout()
        check(self.lib.mcd_scat_layout_get(self.ctx, C.byref(lay)), self.ctx)
        self.row_words = int(lay.row_words)
        self.word_of = list(lay.word_of)
        self.tables = None
        self._order_enqueued = False
        self.cap_el = self.cap_sc = 0
        self.next_id = 0
        self._pinned = {}

    @property
    def channel_names(self):
        """The `type` labels in sigma-dict insertion order; the elastic channel is 'Mott' for elastic = 'Mott'
        (scattering.py:493, :589)."""
        names = list(self.material.channel_names)
        if self.params.elastic == ELASTIC_MOTT:
            names[0] = 'Mott'
        return names

    # ---- lifetime -----------------------------------------------------------------------------
    def clone(self):
        """A second context sharing this engine's tables (mcd_clone) with its own outputs and statistics."""
        import copy
        other = copy.copy(self)
        other.ctx = C.c_void_p()
        check(self.lib.mcd_clone(self.ctx, C.byref(other.ctx)), self.ctx)
        other.cap_el = other.cap_sc = 0
        other._pinned = {}
        other._parent = self                      # keep the owner of the tables alive
        other._clones = []
        other._pipe = None
        self._clones = getattr(self, '_clones', []) + [other]
        return other

    def close(self):
        self._pipe = None
        for c in getattr(self, '_clones', []):    # clones first: they borrow this context's tables
            c.close()
        self._clones = []
        if getattr(self, 'ctx', None) is not None and self.ctx.value:
            self.lib.mcd_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        check(rc, self.ctx)

    # ---- tables -------------------------------------------------------------------------------
    def tuning(self, **kw):
        """Launch tuning of this context (mcd_tuning: handover_passes, handover_threshold, blocks_per_sm,
        statistics_kernel); with keyword arguments the named fields are changed first (before set_capacity).  No
        result depends on them."""
        t = _lib.Tuning()
        self._ck(self.lib.mcd_tuning_get(self.ctx, C.byref(t)))
        if kw:
            for k, v in kw.items():
                if not hasattr(t, k):
                    raise KeyError(k)
                setattr(t, k, int(v))
            self._ck(self.lib.mcd_tuning_set(self.ctx, C.byref(t)))
        return {k: int(getattr(t, k)) for k, _ in _lib.Tuning._fields_}

    def table_build(self, slot, Es, Ws, shell_E=0.0, shell_n=0.0, stream=None):
        Es, Ws = f32(Es), f32(Ws)
        self._ck(self.lib.mcd_table_build(self.ctx, slot, fptr(Es), len(Es), fptr(Ws), len(Ws), shell_E, shell_n,
                                          C.c_void_p(stream or 0)))

    def table_upload(self, slot, Es, Ws, cdf):
        Es, Ws, cdf = f32(Es), f32(Ws), f32(cdf)
        assert cdf.shape == (len(Es), len(Ws) - 1), 'cdf must be nE x (nW-1)'
        self._ck(self.lib.mcd_table_upload(self.ctx, slot, fptr(Es), len(Es), fptr(Ws), len(Ws), fptr(cdf)))

    def mott_upload(self, Es, sigmas, probTable, thetas):
        """The table of elastic = 'Mott' (mottTable of extFunctions.py: Es ascending, sigmas in cm^2, probTable
        nE x nTheta, thetas in degrees; float64 like the reference keeps them)."""
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        Es, sigmas, probTable, thetas = f64(Es), f64(sigmas), f64(probTable), f64(thetas)
        assert probTable.shape == (len(Es), len(thetas)) and len(sigmas) == len(Es)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self._ck(self.lib.mcd_mott_upload(self.ctx, len(Es), len(thetas), dp(Es), dp(sigmas), dp(probTable), dp(thetas)))

    def table_download(self, slot):
        nE, nW = C.c_uint32(), C.c_uint32()
        self._ck(self.lib.mcd_table_shape(self.ctx, slot, C.byref(nE), C.byref(nW)))
        out = np.empty((nE.value, nW.value - 1), dtype=np.float32)
        self._ck(self.lib.mcd_table_download(self.ctx, slot, fptr(out)))
        return out

    # ---- runs ---------------------------------------------------------------------------------
    def set_capacity(self, max_electrons, max_scat_rows=0):
        if max_electrons != self.cap_el or max_scat_rows != self.cap_sc:
            self._ck(self.lib.mcd_set_capacity(self.ctx, int(max_electrons), int(max_scat_rows)))
            self.cap_el, self.cap_sc = max_electrons, max_scat_rows

    def _host(self, key, nbytes, pinned):
        """Host destination for a drain: a fresh array, or with pinned=True a REUSED page-locked staging
        buffer (the returned rows are then only valid until the next drain)."""
        if not pinned:
            return np.empty(max(nbytes, 1), dtype=np.uint8)[:nbytes]
        buf = self._pinned.get(key)
        if buf is None or buf.nbytes < nbytes:
            import torch
            # page-locking is slow (~1-2 GB/s): over-allocate by 12 % so that batches of slightly different row
            # counts reuse the same staging buffer
            self._pinned.pop(key, None)
            self._pinned.pop(key + '_t', None)
            t = torch.empty(max(int(nbytes * 1.125) + 4096, 1), dtype=torch.uint8, pin_memory=True)
            buf = t.numpy()
            self._pinned[key + '_t'] = t
            self._pinned[key] = buf
        return buf[:nbytes]

    def launch(self, first_id, count, stream=None):
        """Asynchronous mcd_run on `stream` (a raw cudaStream_t integer; None = default stream)."""
        self._order_enqueued = False               # mcd_run invalidates any regrouping of the previous run
        self._ck(self.lib.mcd_run(self.ctx, int(first_id), int(count), C.c_void_p(stream or 0)))

    def sync(self, stream=None):
        self._ck(self.lib.mcd_sync(self.ctx, C.c_void_p(stream or 0)))

    def kernel_ms(self):
        ms = C.c_float()
        self._ck(self.lib.mcd_last_kernel_ms(self.ctx, C.byref(ms)))
        return ms.value

    def reset_stats(self, stream=None):
        self._ck(self.lib.mcd_reset_stats(self.ctx, C.c_void_p(stream or 0)))

    def counters(self):
        c = _lib.Counters()
        self._ck(self.lib.mcd_counters_get(self.ctx, C.byref(c)))
        d = {k: int(getattr(c, k)) for k in ('electrons_done', 'events', 'double_records', 'scat_rows_stored',
                                              'scat_rows_dropped', 'nan_clamped', 'select_exact')}
        d['outcome'] = {n: int(c.outcome[i]) for i, n in enumerate(_lib.OUTCOMES)}
        nch = self.cmat.n_shells + 3
        d['channel'] = [int(c.channel[i]) for i in range(nch)]
        return d

    def histograms(self):
        """Backscatter histograms (int64): final energy, exit polar angle from the surface normal, exit
        azimuth, last interior depth; plus bin edges."""
        h = np.empty(_lib.HIST_TOTAL, dtype=np.int64)
        self._ck(self.lib.mcd_hist_get(self.ctx, h.ctypes.data_as(C.POINTER(C.c_int64))))
        return self.split_hist(h)

    def split_hist(self, h):
        lo, hi, dm = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.lib.mcd_hist_ranges(self.ctx, C.byref(lo), C.byref(hi), C.byref(dm)))
        a, b, c_ = _lib.HIST_E_BINS, _lib.HIST_POLAR_BINS, _lib.HIST_AZ_BINS
        return {'energy': (h[:a], np.linspace(lo.value, hi.value, a + 1)),
                'polar_deg': (h[a:a + b], np.linspace(0., 90., b + 1)),
                'azimuth_deg': (h[a + b:a + b + c_], np.linspace(-180., 180., c_ + 1)),
                'depth': (h[a + b + c_:], np.linspace(0., dm.value, _lib.HIST_DEPTH_BINS + 1))}

    def stats_tensor(self):
        """The device int64 histogram||counter block as a torch tensor (zero-copy), for the NCCL
        all-reduce that is the path's only collective."""
        import torch
        ptr, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.mcd_stats_device_ptr(self.ctx, C.byref(ptr), C.byref(n)))

        class _Wrap:
            __cuda_array_interface__ = {'shape': (int(n.value),), 'typestr': '<i8', 'data': (int(ptr.value), False),
                                        'version': 2}
        return torch.as_tensor(_Wrap(), device='cuda:%d' % self.device)

    def launch_info(self):
        b, t, r, o = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.mcd_launch_info(self.ctx, C.byref(b), C.byref(t), C.byref(r), C.byref(o)))
        n = C.c_int()
        self._ck(self.lib.mcd_last_launches(self.ctx, C.byref(n)))
        return {'blocks': b.value, 'threads': t.value, 'regs': r.value, 'blocks_per_sm': o.value,
                'launches_per_run': n.value}

    def enqueue_els_drain(self, count, stream=None):
        """Enqueue the D2H copy of the els rows of the run just launched on `stream`, behind its kernels, into this
        context's pinned staging buffer: the copy starts the moment the run ends and overlaps whatever the other
        contexts of a pipeline are doing.  Returns the (not yet valid) rows; sync(stream) makes them valid."""
        raw = self._host('els', count * 32, True)
        self._ck(self.lib.mcd_drain_els(self.ctx, raw.ctypes.data_as(C.c_void_p), 0, count, C.c_void_p(stream or 0)))
        return raw.view(_lib.EL_ROW_DTYPE)

    def drain(self, first_id, count, want_els=True, want_scats=True, pinned=False, stream=None, ordered=False):
        """Copy the rows of the last run to host arrays.  ordered=True regroups the scats rows on the device
        into trajectory order first (mcd_order_scats) and returns them with per-electron offsets;
        ordered='columns' does the same regrouping into one contiguous array per column (ScatCols) and leaves
        the id / event words on the device (17 instead of 24 bytes per `type, position, energy` row)."""
        st = C.c_void_p(stream or 0)
        n_els, n_sc = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.mcd_row_counts(self.ctx, C.byref(n_els), C.byref(n_sc)))
        els = None
        if want_els:
            raw = self._host('els', count * 32, pinned)
            self._ck(self.lib.mcd_drain_els(self.ctx, raw.ctypes.data_as(C.c_void_p), 0, count, st))
            els = raw.view(_lib.EL_ROW_DTYPE)
        scats = {}
        nsc = int(n_sc.value)
        words = None
        if want_scats and self.row_words and self.cap_sc:
            offs = None
            if nsc and ordered == 'columns':
                if not self._order_enqueued:
                    self._ck(self.lib.mcd_order_scats_columns(self.ctx, st))
                self._order_enqueued = False
                lay = [C.c_uint64() for _ in range(5)]
                self._ck(self.lib.mcd_scat_columns_layout(self.ctx, *[C.byref(v) for v in lay]))
                n_rows, _, stride, type_off, total = [int(v.value) for v in lay]
                block = self._host('scats', total, pinned)
                self._ck(self.lib.mcd_drain_scats_columns(self.ctx, block.ctypes.data_as(C.c_void_p), total, st))
                offs = self._host('offsets', (count + 1) * 8, pinned).view(np.uint64)
                self._ck(self.lib.mcd_scat_offsets(self.ctx, offs.ctypes.data_as(C.POINTER(C.c_uint64)), 0, count, st))
                self.sync(stream)
                return els, ScatCols(block, n_rows, stride, type_off, self.word_of, first_id, offs)
            rawb = self._host('scats', nsc * self.row_words * 4, pinned)
            if nsc and ordered:
                self._ck(self.lib.mcd_order_scats(self.ctx, st))
                self._ck(self.lib.mcd_drain_scats_ordered(self.ctx, rawb.ctypes.data_as(C.c_void_p), 0, nsc, st))
                offs = self._host('offsets', (count + 1) * 8, pinned).view(np.uint64)
                self._ck(self.lib.mcd_scat_offsets(self.ctx, offs.ctypes.data_as(C.POINTER(C.c_uint64)), 0, count, st))
            elif nsc:
                self._ck(self.lib.mcd_drain_scats(self.ctx, rawb.ctypes.data_as(C.c_void_p), 0, nsc, st))
            words = rawb.view(np.uint32).reshape(nsc, self.row_words)
        self.sync(stream)
        if words is not None:
            scats = ScatRows(words, self.word_of, first_id, offs)
        return els, scats

    def run(self, count=None, first_id=None, max_scat_rows=None, pinned=False, stream=None, drain=True, ordered=False):
        """Simulate `count` electrons (default inputPar['num_el']) with ids starting at `first_id`
        (default: continue after the previous run) and return a RunResult."""
        count = int(count if count is not None else self.inputPar['num_el'])
        first_id = int(self.next_id if first_id is None else first_id)
        if count <= 0:                                 # nothing to do (the reference spawns no job for num_el = 0)
            empty = np.empty(0, _lib.EL_ROW_DTYPE)
            zero = {k: 0 for k in ('electrons_done', 'events', 'double_records', 'scat_rows_stored',
                                   'scat_rows_dropped', 'nan_clamped', 'select_exact')}
            zero['outcome'] = {n: 0 for n in _lib.OUTCOMES}
            zero['channel'] = [0] * (self.cmat.n_shells + 3)
            return RunResult(first_id, 0, empty, {}, zero, 0.0)
        if self.params.mode == _lib.MODE_DS and self.tables is None:
            from .probTables import genTables
            genTables(self.inputPar, engine=self)
        auto = max_scat_rows is None
        if auto:
            max_scat_rows = 0
            if self.row_words:
                # a trajectory has at most maxScatt events; typical means are a few hundred
                max_scat_rows = count * min(int(self.inputPar['maxScatt']), 512) + 65536
        while True:
            self.set_capacity(max(count, self.cap_el), max(max_scat_rows, self.cap_sc))
            self.reset_stats(stream)
            self.launch(first_id, count, stream)
            self.sync(stream)
            cnt = self.counters()
            if not cnt['scat_rows_dropped']:
                break
            # Philox streams are keyed on the electron id, so re-running the batch reproduces it exactly
            if not auto or self.cap_sc >= count * int(self.inputPar['maxScatt']):
                raise _lib.MCDError(-5, '%d scats rows did not fit (capacity %d): raise max_scat_rows or run '
                                    'fewer electrons per batch' % (cnt['scat_rows_dropped'], self.cap_sc))
            max_scat_rows = min(2 * self.cap_sc, count * int(self.inputPar['maxScatt']))
        self.next_id = max(self.next_id, first_id + count)
        els, scats = self.drain(first_id, count, pinned=pinned, stream=stream, ordered=ordered) if drain else (None, {})
        return RunResult(first_id, count, els, scats, cnt, self.kernel_ms())

    def pipeline_engines(self):
        """This engine and the clones run_batches rotates over."""
        pipe = getattr(self, '_pipe', None)
        return [self] + (list(pipe[0]) if pipe else [])

    def stats_host(self):
        """The int64 histogram||counter block of this context, copied to the host (numpy)."""
        return self.stats_tensor().cpu().numpy()

    def merge_stats_from(self, other, stream=None):
        """This context's histogram||counter block += `other`'s (which is cleared), on `stream` (mcd_stats_merge)."""
        self._ck(self.lib.mcd_stats_merge(self.ctx, other.ctx, C.c_void_p(stream or 0)))

    def pipeline_stats(self, device=True):
        """The histogram||counter block of a run made of run_batches(..., accumulate=True) calls: what is all-reduced
        ONCE at the end of the run.  Statistics-only contexts accumulate on the device over their
        batches; the clones' blocks are merged into this context's block (mcd_stats_merge) and the returned tensor IS
        that block (zero-copy, all-reduce it in place).  When scats rows are recorded the per-batch blocks were summed
        by the host (run_batches) and are uploaded.  device=False returns a numpy copy instead."""
        import torch
        if self.row_words:
            acc = getattr(self, '_acc_host', None)
            tot = acc if acc is not None else np.zeros(_lib.HIST_TOTAL + C.sizeof(_lib.Counters) // 8, np.int64)
            return torch.from_numpy(tot.copy()).to('cuda:%d' % self.device) if device else tot.copy()
        torch.cuda.synchronize(self.device)              # every stream of the pipeline has finished its batches
        st = torch.cuda.current_stream(self.device).cuda_stream
        for e in self.pipeline_engines()[1:]:
            self.merge_stats_from(e, st)
        t = self.stats_tensor()
        return t if device else t.cpu().numpy()

    def reset_pipeline_stats(self):
        for e in self.pipeline_engines():
            e.reset_stats()
            e.sync()
        self._acc_host = None

    def run_batches(self, total, batch, first_id=0, max_scat_rows=None, on_launched=None, ordered=False, depth=None,
                    accumulate=False, els=None):
        """Simulate `total` electrons in batches of `batch`, PIPELINED over `depth` contexts / CUDA streams
        (default 6 for statistics-only runs, 2 when scats rows are recorded): while the host drains batch k
        (D2H into pinned buffers) the next batches already run, their blocks fill the SMs that the tail of batch k
        leaves idle, and the tail hand-over passes of one batch hide behind the main pass of another.  Yields
        one RunResult per batch, in order.  Rows of a yielded result live in reused pinned buffers: consume them
        before advancing the generator `depth` times.
        `on_launched(engine, torch_stream)` is called right after each launch.
        accumulate=True: the statistics of the batches are summed for ONE all-reduce at the end of the run
        (pipeline_stats()): statistics-only contexts simply keep accumulating on the device (no per-batch reset, no
        per-batch counter read; RunResult.counters is None); when scats rows are recorded the per-batch block is read
        with the batch's counters and summed by the host.
        els=False: the els rows stay on the device (histogram-only runs, BASELINE C4)."""
        import torch
        from collections import deque
        if self.params.mode == _lib.MODE_DS and self.tables is None:
            from .probTables import genTables
            genTables(self.inputPar, engine=self)
        depth = int(depth or (2 if self.row_words else 6))
        # the extra contexts, their device buffers, the pinned staging buffers and the streams are kept for
        # later calls (creating them costs tens of milliseconds of cudaMalloc / page-locking)
        pipe = getattr(self, '_pipe', None)
        if pipe is None:
            pipe = self._pipe = ([], [])
        while len(pipe[0]) < depth - 1:
            pipe[0].append(self.clone())
        while len(pipe[1]) < depth:
            pipe[1].append(torch.cuda.Stream(device=self.device))
        engs = [self] + pipe[0][:depth - 1]
        streams = pipe[1][:depth]
        if max_scat_rows is None:
            max_scat_rows = (batch * min(int(self.inputPar['maxScatt']), 512) + 65536) if self.row_words else 0
        pending = deque()

        want_els = bool(self.el_cols) if els is None else bool(els)

        def finish(p):
            e, st, first, cnt, early = p
            e.sync(st)
            cnts = None
            if e.row_words or not accumulate:
                cnts = e.counters()
                if cnts['scat_rows_dropped']:
                    raise _lib.MCDError(-5, '%d scats rows did not fit: raise max_scat_rows or lower batch' % cnts['scat_rows_dropped'])
                if accumulate:
                    blk = e.stats_host()
                    self._acc_host = blk if getattr(self, '_acc_host', None) is None else self._acc_host + blk
            if early is not None:                          # the copy was enqueued behind the kernel: it is done
                els_rows, scats = early, {}
            elif want_els or e.row_words:
                els_rows, scats = e.drain(first, cnt, want_els=want_els, pinned=True, stream=st, ordered=ordered)
            else:
                els_rows, scats = None, {}
            return RunResult(first, cnt, els_rows, scats, cnts, e.kernel_ms())
        k = 0
        try:
            for first in range(first_id, first_id + total, batch):
                cnt = min(batch, first_id + total - first)
                e, st = engs[k % depth], streams[k % depth].cuda_stream
                e.set_capacity(max(batch, e.cap_el), max(max_scat_rows, e.cap_sc))
                if e.row_words or not accumulate:
                    e.reset_stats(st)
                e.launch(first, cnt, st)
                if ordered == 'columns' and e.row_words:
                    # regrouping enqueued behind the kernel: it overlaps the D2H drain of the previous batch
                    e._ck(e.lib.mcd_order_scats_columns_enqueue(e.ctx, st))
                    e._order_enqueued = True
                if on_launched is not None:
                    with torch.cuda.stream(streams[k % depth]):
                        on_launched(e, streams[k % depth])
                # runs without scats rows: the els rows leave the device right behind the kernel (no host round trip
                # between the end of the run and the start of its copy; at the end of the job only the LAST batch's
                # copy is exposed instead of the copies of all `depth` batches still in flight)
                early = e.enqueue_els_drain(cnt, st) if (want_els and not e.row_words) else None
                pending.append((e, st, first, cnt, early))
                if len(pending) == depth:                  # every context is busy: drain the oldest batch
                    yield finish(pending.popleft())
                k += 1
            while pending:
                yield finish(pending.popleft())
            self.next_id = max(self.next_id, first_id + total)
        finally:
            for st in streams:
                st.synchronize()

    # ---- unit hooks (parity tests) --------------------------------------------------------------
    def test_sigmas(self, E):
        E = f32(E)
        nch = self.cmat.n_shells + 3
        sig = np.empty((len(E), nch), np.float32)
        mfp = np.empty(len(E), np.float32)
        self._ck(self.lib.mcd_test_sigmas(self.ctx, fptr(E), len(E), fptr(sig), fptr(mfp)))
        return sig, mfp

    def test_cum(self, E):
        """(cum[n, S+3] in slot order, slot->channel map, uses_poly, verified poly error)"""
        E = f32(E)
        nch = self.cmat.n_shells + 3
        cum = np.empty((len(E), nch), np.float32)
        sl = np.empty(nch, np.int32)
        up, err = C.c_int32(), C.c_double()
        self._ck(self.lib.mcd_test_cum(self.ctx, fptr(E), len(E), fptr(cum), iptr(sl), C.byref(up), C.byref(err)))
        return cum, sl, bool(up.value), err.value

    def test_select(self, sigma, u):
        sigma, u = f32(sigma), f32(u)
        pick = np.empty(len(u), np.int32)
        path = np.empty(len(u), np.int32)
        self._ck(self.lib.mcd_test_select(self.ctx, fptr(sigma), sigma.shape[1], fptr(u), len(u), iptr(pick), iptr(path)))
        return pick, path

    def test_ruth(self, E, u):
        E, u = f32(E), f32(u)
        c2, om = np.empty(len(E), np.float32), np.empty(len(E), np.float32)
        self._ck(self.lib.mcd_test_ruth(self.ctx, fptr(E), fptr(u), len(E), fptr(c2), fptr(om)))
        return c2, om

    def test_bincoll(self, E, W):
        E, W = f32(E), f32(W)
        c2, om = np.empty(len(E), np.float32), np.empty(len(E), np.float32)
        self._ck(self.lib.mcd_test_bincoll(self.ctx, fptr(E), fptr(W), len(E), fptr(c2), fptr(om)))
        return c2, om

    def test_rotate(self, d, y, c2, om, u_phi):
        d, y, c2, om, u_phi = f32(d), f32(y), f32(c2), f32(om), f32(u_phi)
        do, yo = np.empty_like(d), np.empty_like(y)
        self._ck(self.lib.mcd_test_rotate(self.ctx, fptr(d), fptr(y), fptr(c2), fptr(om), fptr(u_phi), len(c2),
                                          fptr(do), fptr(yo)))
        return do, yo

    def test_stopping(self, model, E):
        E = f32(E)
        sp = np.empty(len(E), np.float32)
        self._ck(self.lib.mcd_test_stopping(self.ctx, _lib.BETHE[model], fptr(E), len(E), fptr(sp)))
        return sp

    def test_lookup(self, slot, E, u):
        E, u = f32(E), f32(u)
        W = np.empty(len(E), np.float32)
        iE, jW = np.empty(len(E), np.int32), np.empty(len(E), np.int32)
        self._ck(self.lib.mcd_test_lookup(self.ctx, slot, fptr(E), fptr(u), len(E), fptr(W), iptr(iE), iptr(jW)))
        return W, iE, jW

    def test_mott(self, E, u):
        """(sigma in cm^2, cos^2(theta/2), iE, jTheta) of the uploaded Mott table for energies E and uniforms u."""
        E, u = f32(E), f32(u)
        sg, c2 = np.empty(len(E), np.float32), np.empty(len(E), np.float32)
        iE, jT = np.empty(len(E), np.int32), np.empty(len(E), np.int32)
        self._ck(self.lib.mcd_test_mott(self.ctx, fptr(E), fptr(u), len(E), fptr(sg), fptr(c2), iptr(iE), iptr(jT)))
        return sg, c2, iE, jT

    def test_step(self, state, u):
        state, u = f32(state), f32(u)
        so = np.empty_like(state)
        row = np.empty((len(state), 8), np.float32)
        self._ck(self.lib.mcd_test_step(self.ctx, fptr(state), fptr(u), len(state), fptr(so), fptr(row)))
        return so, row


def philox(seed, el_id, ev):
    lib = _lib.load()
    raw = (C.c_uint32 * 4)()
    u = (C.c_float * 4)()
    check(lib.mcd_test_philox(int(seed), int(el_id), int(ev), raw, u))
    return list(raw), list(u)
