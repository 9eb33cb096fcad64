"""Result store with the reference's layout (MC/doScatter.py:63-67, 99-119, 175-204): three keys
  'input'  the input parameters as strings,
  'els'    one row per saveOutcomes call, columns in zipDict order (fileTools.py:67-99),
  'scats'  one row per scattering event.

Two back ends with the same keys, column names, order and dtypes (float64 / str):
  * HDF5 through pandas.HDFStore (PyTables 'table' format, data_columns=True, appendable) -- exactly what the
    reference writes -- when PyTables is importable;
  * a directory of Parquet files (pyarrow) otherwise.  This image ships no HDF5 library (no PyTables, h5py or
    libhdf5), so Parquet is what runs here; `fmt='auto'` picks HDF5 wherever it is available.
Rows are appended in batches straight from the columnar GPU output: the per-4-electron open/append/close cycle
that made the reference lock up above 5e5 electrons (README.md:58-59) does not exist.
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


class ResultStore:
    """Appendable store; use as a context manager."""

    def __init__(self, path, inputPar, fmt='auto'):
        import pandas as pd
        self.inputPar = dict(inputPar)
        if fmt == 'auto':
            fmt = 'hdf5' if have_pytables() else 'parquet'
        self.fmt = fmt
        self.path = path
        self._pq = {}
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
            raise ValueError('fmt must be auto, hdf5 or parquet')

    def append(self, key, df):
        if len(df) == 0:
            return
        if self.fmt == 'hdf5':
            # widths of the longest labels (the reference's min_itemsize=6 truncates 'Gryz2s2p', doScatter.py:75-79)
            sizes = {c: 16 for c in df.columns if df[c].dtype == object}
            self.store.put(key, df, format='table', data_columns=True, append=True, min_itemsize=sizes)
        else:
            import pyarrow as pa
            import pyarrow.parquet as pq
            tab = pa.Table.from_pandas(df, preserve_index=False)
            w = self._pq.get(key)
            if w is None:
                w = self._pq[key] = pq.ParquetWriter(os.path.join(self.path, key + '.parquet'), tab.schema)
            w.write_table(tab)

    def append_result(self, res, channel_names, double_record=False):
        self.append('els', els_frame(res, self.inputPar, double_record))
        if len(res.scats):
            self.append('scats', scats_frame(res, self.inputPar, channel_names))

    def close(self):
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


def read_store(path):
    """Read back {'input': Series/dict, 'els': DataFrame, 'scats': DataFrame or None} from either back end."""
    import pandas as pd
    if os.path.isdir(path):
        out = {'input': json.load(open(os.path.join(path, 'input.json')))}
        for key in ('els', 'scats'):
            f = os.path.join(path, key + '.parquet')
            out[key] = pd.read_parquet(f) if os.path.exists(f) else None
        return out
    with pd.HDFStore(path, mode='r') as st:
        return {'input': st['input'], 'els': st['els'], 'scats': st['scats'] if '/scats' in st.keys() else None}
