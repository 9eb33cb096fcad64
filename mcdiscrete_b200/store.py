"""Result store with the reference's layout (MC/doScatter.py:63-67, 99-119, 175-204): three keys
  'input'  the input parameters as strings,
  'els'    one row per saveOutcomes call, columns in zipDict order (fileTools.py:67-99),
  'scats'  one row per scattering event.

Three back ends with the same keys, column names, order and dtypes (float64 / str):
  * 'hdf5' with PyTables importable: pandas.HDFStore (PyTables 'table' format, data_columns=True, appendable) --
    exactly what the reference writes;
  * 'hdf5' without PyTables (this image ships no HDF5 library at all): the self-contained writer
    mcdiscrete_b200/h5lite.py produces the same file layout -- groups /input, /els, /scats, each holding a chunked,
    unlimited-length compound dataset `table` (int64 `index` + one field per column) with the PyTables and pandas
    attributes of an appendable frame table.  Columns go from the GPU's arrays straight into the compound rows (no
    DataFrame, no per-row Python objects); chunks stream to disk as batches arrive;
  * 'parquet': a directory of Parquet files (pyarrow).
`fmt='auto'` = 'hdf5'.  Rows are appended in batches straight from the columnar GPU output: the per-4-electron
open/append/close cycle that made the reference lock up above 5e5 electrons (README.md:58-59) does not exist.
"""
import json
import os

import numpy as np

from . import _lib
from .fileTools import flat_columns


def store_name(inputPar, storePath='../data/', ext='.h5'):
    """doScatter.py:175-186."""
    return (storePath + 'TRSM_diff:' + str(inputPar.get('diffMFP', False)) + '_thick:' + str(inputPar['thickness']) +
            '_mat:' + str(inputPar['material']) + '_mode:' + str(inputPar['mode']) +
            '_elastic:' + str(inputPar.get('elastic', 'Ruth_vanilla')) + '_tilt:' + str(inputPar['s_tilt']) +
            '_Emin:' + str(inputPar['Emin']) + '_E0:' + str(inputPar['E0']) + '_tolE:' + str(inputPar.get('tol_E')) +
            '_tolW:' + str(inputPar.get('tol_W')) + ext)


def have_pytables():
    try:
        import tables  # noqa: F401
        return True
    except Exception:
        return False


def els_frame(res, inputPar, double_record=False):
    """els rows of a RunResult as a pandas DataFrame in the reference's column order."""
    import pandas as pd
    cols = flat_columns(inputPar['electron_output'], 'els')
    e = res.els
    idx = np.arange(len(e))
    if double_record:                      # the reference appends a second 'abs' row
        idx = np.sort(np.concatenate([idx, idx[(e['code'] & 0x100) != 0]]))
        second = np.zeros(len(idx), bool)
        second[1:] = idx[1:] == idx[:-1]
    data = {}
    for c in cols:
        if c == 'outcome':
            names = np.array(_lib.OUTCOMES, dtype=object)[(e['code'][idx] & 0xff)]
            if double_record:
                names = names.copy()
                names[second] = 'abs'
            data[c] = names
        elif c == 'final_E':
            data[c] = e['final_E'][idx].astype(np.float64)
        else:
            data[c] = e[c][idx].astype(np.float64)
    return pd.DataFrame(data, columns=cols)


def scats_frame(res, inputPar, channel_names, sort=True):
    """scats rows as a DataFrame: scalar columns in requested order, then x, y, z (zipDict order).  'direction'
    stays one object column in the reference (electron.py:99); here it is three float columns dir_x/dir_y/dir_z."""
    import pandas as pd
    sc = res.sorted_scats() if sort else res.scats
    cont = 'DS' not in inputPar['mode']
    req = [c for c in inputPar['scatter_output'] if not (cont and c in ('position', 'E'))]
    cols = flat_columns(req, 'scats')
    data = {}
    out_cols = []
    for c in cols:
        if c == 'type':
            if cont:
                data[c] = np.full(len(sc), inputPar.get('elastic', 'Ruth_vanilla'), dtype=object)
            else:
                data[c] = np.array(channel_names, dtype=object)[sc['type']]
            out_cols.append(c)
        elif c == 'direction':
            d = sc['direction']
            for k, nm in enumerate(('dir_x', 'dir_y', 'dir_z')):
                data[nm] = d[:, k].astype(np.float64)
                out_cols.append(nm)
        else:
            data[c] = np.asarray(sc[c], dtype=np.float64)
            out_cols.append(c)
    return pd.DataFrame(data, columns=out_cols)


def _label_width():
    """Width of the string columns: the longest label they can hold -- outcome names, channel names of every
    material ('Gryz3s3p3d') and, in continuous mode, the elastic label ('Ruth_vanilla_wDefl', 18 characters).  The
    reference's min_itemsize = 6 truncates most of them."""
    from .engine import ELASTIC_MODELS
    from .material import material, _SPECIES as SPECIES
    w = max(len(x) for x in tuple(_lib.OUTCOMES) + tuple(ELASTIC_MODELS))
    for sp in SPECIES:
        w = max([w] + [len(x) for x in material(sp).channel_names])
    return max(w, 18)


STR_WIDTH = _label_width()


def _as_label(value, n):
    """n copies of a label as fixed-width bytes; refuses to truncate."""
    b = str(value).encode()
    if len(b) > STR_WIDTH:
        raise ValueError('label %r is longer than the %d-byte string column' % (value, STR_WIDTH))
    return np.full(n, b, dtype='S%d' % STR_WIDTH)


def els_columns(res, inputPar, double_record=False):
    """els rows as {column: array} in the reference's column order; 'outcome' as fixed-width byte strings."""
    cols = flat_columns(inputPar['electron_output'], 'els')
    e = res.els
    idx = np.arange(len(e))
    second = None
    if double_record:                      # the reference appends a second 'abs' row
        idx = np.sort(np.concatenate([idx, idx[(e['code'] & 0x100) != 0]]))
        second = np.zeros(len(idx), bool)
        second[1:] = idx[1:] == idx[:-1]
    out = {}
    for c in cols:
        if c == 'outcome':
            names = np.array(_lib.OUTCOMES, dtype='S%d' % STR_WIDTH)[(e['code'][idx] & 0xff)]
            if second is not None:
                names[second] = b'abs'
            out[c] = names
        else:
            out[c] = e[c][idx].astype(np.float64)
    return out


def scats_columns(res, inputPar, channel_names, sort=True):
    """scats rows as {column: array}: scalar columns in requested order, then x, y, z (zipDict order)."""
    sc = res.sorted_scats() if sort else res.scats
    cont = 'DS' not in inputPar['mode']
    req = [c for c in inputPar['scatter_output'] if not (cont and c in ('position', 'E'))]
    out = {}
    for c in flat_columns(req, 'scats'):
        if c == 'type':
            if cont:
                out[c] = _as_label(inputPar.get('elastic', 'Ruth_vanilla'), len(sc))
            else:
                out[c] = np.array(channel_names, dtype='S%d' % STR_WIDTH)[sc['type']]
        elif c == 'direction':
            d = sc['direction']
            for k, nm in enumerate(('dir_x', 'dir_y', 'dir_z')):
                out[nm] = d[:, k].astype(np.float64)
        else:
            out[c] = np.asarray(sc[c], dtype=np.float64)
    return out


class _H5LiteTable:
    """One pandas-style appendable frame table (`/key/table`) written through h5lite."""

    def __init__(self, writer, key, cols, chunk_rows):
        self.w, self.key = writer, key
        self.names = list(cols)
        self.dtype = np.dtype([('index', '<i8')] + [(n, cols[n].dtype.newbyteorder('<') if cols[n].dtype.kind != 'S'
                                                     else cols[n].dtype) for n in self.names])
        self.table = writer.create_table('/%s/table' % key, self.dtype, chunk_rows)
        self.rows = 0

    def append(self, cols):
        n = len(next(iter(cols.values()))) if cols else 0
        if list(cols) != self.names:
            raise ValueError('columns %s do not match the table %s' % (list(cols), self.names))
        step = 1 << 20                       # build the compound rows in slices: bounded extra host memory
        for a in range(0, n, step):
            b = min(n, a + step)
            rows = np.empty(b - a, self.dtype)
            rows['index'] = np.arange(self.rows + a, self.rows + b)
            for k in self.names:
                rows[k] = cols[k][a:b]
            self.w.append(self.table, rows)
        self.rows += n

    def finish(self):
        names = self.names
        self.w.set_attrs('/' + self.key, {
            'CLASS': 'GROUP', 'TITLE': '', 'VERSION': '1.0', 'pandas_type': 'frame_table',
            'pandas_version': '0.15.2', 'table_type': 'appendable_frame', 'index_cols': [(0, 'index')],
            'values_cols': list(names), 'non_index_axes': [(1, list(names))], 'data_columns': list(names),
            'nan_rep': 'nan', 'encoding': 'UTF-8', 'errors': 'strict', 'levels': 1, 'metadata': [],
            'info': {1: {'type': 'Index', 'names': [None]}, 'index': {}}})
        at = {'CLASS': 'TABLE', 'VERSION': '2.7', 'TITLE': '', 'NROWS': np.int64(self.rows), 'index_kind': 'integer'}
        for i, n in enumerate(self.dtype.names):
            sub = self.dtype.fields[n][0]
            at['FIELD_%d_NAME' % i] = n
            at['FIELD_%d_FILL' % i] = b'' if sub.kind == 'S' else np.zeros((), sub)[()]
        for n in names:
            sub = self.dtype.fields[n][0]
            at[n + '_kind'] = [n]
            at[n + '_meta'] = None
            at[n + '_dtype'] = sub.name              # 'float64', 'bytes128'
        self.w.set_attrs('/%s/table' % self.key, at)


class ResultStore:
    """Appendable store; use as a context manager."""

    def __init__(self, path, inputPar, fmt='auto', chunk_rows=1 << 16):
        self.inputPar = dict(inputPar)
        if fmt == 'auto':
            fmt = 'hdf5'
        self.fmt = fmt
        self.path = path
        self._pq = {}
        self._h5 = None
        if fmt == 'h5lite' or (fmt == 'hdf5' and not have_pytables()):      # 'h5lite' forces the self-contained writer
            from . import h5lite
            if os.path.exists(path):
                os.remove(path)
            self.fmt = 'h5lite'
            self._h5 = h5lite.Writer(path)
            self._h5.set_attrs('/', {'CLASS': 'GROUP', 'PYTABLES_FORMAT_VERSION': '2.1', 'TITLE': '', 'VERSION': '1.0'})
            self._h5_tables = {}
            self._chunk_rows = int(chunk_rows)
            # doScatter.py:201-203: store['input'] = one-row frame of the parameters (here: as strings)
            vals = {str(k): np.array([str(v).encode()]) for k, v in self.inputPar.items()}
            t = _H5LiteTable(self._h5, 'input', vals, 1)
            t.append(vals)
            self._h5_tables['input'] = t
            return
        import pandas as pd
        ser = pd.Series([str(v) for v in self.inputPar.values()], index=list(self.inputPar.keys()), dtype=str)
        if fmt == 'hdf5':
            if os.path.exists(path):
                os.remove(path)
            self.store = pd.HDFStore(path, mode='w')
            self.store['input'] = ser                                  # doScatter.py:201-203
        elif fmt == 'parquet':
            os.makedirs(path, exist_ok=True)
            for f in os.listdir(path):
                if f.endswith('.parquet') or f == 'input.json':
                    os.remove(os.path.join(path, f))
            json.dump({k: str(v) for k, v in self.inputPar.items()}, open(os.path.join(path, 'input.json'), 'w'), indent=1)
        else:
            raise ValueError('fmt must be auto, hdf5, h5lite or parquet')

    def append(self, key, df):
        if len(df) == 0:
            return
        if self.fmt == 'h5lite':
            def fixed(col):
                b = col.astype(str).str.encode('utf-8')
                if len(b) and int(b.str.len().max()) > STR_WIDTH:
                    raise ValueError('a label of column %r is longer than the %d-byte string column' % (col.name, STR_WIDTH))
                return np.asarray(b, dtype='S%d' % STR_WIDTH)
            self.append_columns(key, {c: (fixed(df[c]) if df[c].dtype == object or str(df[c].dtype).startswith('str')
                                          else np.asarray(df[c])) for c in df.columns})
            return
        if self.fmt == 'hdf5':
            # widths of the longest labels (the reference's min_itemsize=6 truncates 'Gryz2s2p', doScatter.py:75-79)
            sizes = {c: STR_WIDTH for c in df.columns if df[c].dtype == object}
            self.store.put(key, df, format='table', data_columns=True, append=True, min_itemsize=sizes)
        else:
            import pyarrow as pa
            import pyarrow.parquet as pq
            tab = pa.Table.from_pandas(df, preserve_index=False)
            w = self._pq.get(key)
            if w is None:
                w = self._pq[key] = pq.ParquetWriter(os.path.join(self.path, key + '.parquet'), tab.schema)
            w.write_table(tab)

    def append_columns(self, key, cols):
        """h5lite back end: {column: array} -> compound rows of /key/table."""
        t = self._h5_tables.get(key)
        if t is None:
            t = self._h5_tables[key] = _H5LiteTable(self._h5, key, cols, self._chunk_rows)
        t.append(cols)

    def append_result(self, res, channel_names, double_record=False):
        if self.fmt == 'h5lite':
            self.append_columns('els', els_columns(res, self.inputPar, double_record))
            if len(res.scats):
                self.append_columns('scats', scats_columns(res, self.inputPar, channel_names))
            return
        self.append('els', els_frame(res, self.inputPar, double_record))
        if len(res.scats):
            self.append('scats', scats_frame(res, self.inputPar, channel_names))

    def close(self):
        if self.fmt == 'h5lite':
            if self._h5 is not None:
                for t in self._h5_tables.values():
                    t.finish()
                self._h5.close()
                self._h5 = None
            return
        if self.fmt == 'hdf5':
            self.store.close()
        for w in self._pq.values():
            w.close()
        self._pq = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


def read_store(path, reader='auto'):
    """Read back {'input': Series/dict, 'els': DataFrame, 'scats': DataFrame or None} from either back end."""
    import pandas as pd
    if os.path.isdir(path):
        out = {'input': json.load(open(os.path.join(path, 'input.json')))}
        for key in ('els', 'scats'):
            f = os.path.join(path, key + '.parquet')
            out[key] = pd.read_parquet(f) if os.path.exists(f) else None
        return out
    if reader == 'h5lite' or not have_pytables():
        return read_store_h5lite(path)
    with pd.HDFStore(path, mode='r') as st:
        return {'input': st['input'], 'els': st['els'], 'scats': st['scats'] if '/scats' in st.keys() else None}


def read_frame_h5lite(f, key):
    """/key/table of an HDF5 file (h5lite reader) -> DataFrame, the way pandas.read_hdf returns a frame table."""
    import pandas as pd
    g = f[key]
    if g.attr('pandas_type') != 'frame_table':
        raise ValueError('%s is not a pandas frame table' % key)
    rows = g['table'].read()
    cols = g.attr('non_index_axes')[0][1]
    data = {}
    for c in cols:
        v = rows[c]
        data[c] = np.char.decode(v, g.attr('encoding')).astype(object) if v.dtype.kind == 'S' else np.ascontiguousarray(v)
    return pd.DataFrame(data, columns=cols, index=pd.Index(np.ascontiguousarray(rows['index'])))


def read_store_h5lite(path):
    from . import h5lite
    with h5lite.File(path) as f:
        inp = read_frame_h5lite(f, 'input')
        out = {'input': inp.iloc[0], 'els': read_frame_h5lite(f, 'els') if 'els' in f else None}
        out['scats'] = read_frame_h5lite(f, 'scats') if 'scats' in f else None
    return out
