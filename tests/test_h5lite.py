"""CPU-only: the self-contained HDF5 writer / reader (mcdiscrete_b200/h5lite.py).

No HDF5 library exists in this image, so the reader is first checked against a file written by the real libhdf5
(tests/golden/libhdf5_matlab73_testdouble.mat = scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, a MATLAB 7.3
file, BSD-licensed scipy test data) and then serves as the judge of everything the writer produces."""
import os
import pickle
import struct

import numpy as np
import pytest

from mcdiscrete_b200 import h5lite

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_reader_parses_a_libhdf5_written_file():
    f = h5lite.File(os.path.join(GOLD, 'libhdf5_matlab73_testdouble.mat'))
    assert f.base == 512 and f.superblock_version == 0                   # 512-byte MATLAB user block
    assert f.keys() == ['testdouble']
    d = f['testdouble']
    assert not d.is_group and d.shape == (9, 1) and d.dtype == np.dtype('<f8')
    assert d.layout['class'] == 'contiguous'
    assert d.attrs == {'MATLAB_class': b'double'}
    np.testing.assert_allclose(d.read().ravel(), np.arange(9) * np.pi / 4, rtol=1e-15)
    sib = os.path.join(os.path.dirname(__import__('scipy.io').io.__file__), 'matlab', 'tests', 'data',
                       'testdouble_7.4_GLNX86.mat')
    if os.path.exists(sib):                                               # the same variable in MATLAB-5 format
        from scipy.io import loadmat
        assert np.array_equal(loadmat(sib)['testdouble'].ravel(), d.read().ravel())


def test_datatype_messages_match_libhdf5_encodings():
    # float64 / scalar-string encodings as found in the libhdf5-written file above
    assert h5lite.encode_dtype('<f8').hex() == '11203f000800000000004000340b0034ff030000'
    assert h5lite.encode_dtype('S6').hex() == '1300000006000000'
    dt = np.dtype([('index', '<i8'), ('type', 'S16'), ('energy', '<f8'), ('n', '<u2'), ('f', '<f4')])
    back, used, info = h5lite.decode_dtype(h5lite.encode_dtype(dt))
    assert back == dt and used == len(h5lite.encode_dtype(dt)) and info['class'] == 6


def _validate(path):
    """Walk the whole file and check the invariants libhdf5 relies on; returns the number of objects visited."""
    size = os.path.getsize(path)
    f = h5lite.File(path)
    assert f.base == 0 and f.eof == size
    raw = open(path, 'rb').read()
    assert raw[:8] == h5lite.SIGNATURE and len(raw) % 8 == 0
    count = [0]

    def header(addr):
        assert addr % 8 == 0
        ver, _, nmsg, ref, hsize = struct.unpack_from('<BBHII', raw, addr)
        assert ver == 1 and ref == 1 and hsize % 8 == 0
        p, n = addr + 16, 0
        while p < addr + 16 + hsize:
            t, sz, fl = struct.unpack_from('<HHB', raw, p)
            assert sz % 8 == 0
            p += 8 + sz
            n += 1
        assert p == addr + 16 + hsize and n == nmsg

    def group(node):
        header(node.addr)
        count[0] += 1
        bt, hp = struct.unpack_from('<QQ', bytes(node._msg(0x11)), 0)
        assert raw[hp:hp + 4] == b'HEAP' and raw[bt:bt + 4] == b'TREE'
        seg_size, free, seg = struct.unpack_from('<QQQ', raw, hp + 8)
        nxt, fsz = struct.unpack_from('<QQ', raw, seg + free)
        assert nxt == 1 and free + fsz == seg_size                        # one free block up to the end
        typ, level, used = struct.unpack_from('<BBH', raw, bt + 4)
        assert (typ, level) == (0, 0) and 0 <= used <= 32      # libhdf5 too gives an empty group an empty node
        names = []
        for i in range(used):
            key_hi = struct.unpack_from('<Q', raw, bt + 24 + 16 * i + 16)[0]
            child = struct.unpack_from('<Q', raw, bt + 24 + 16 * i + 8)[0]
            assert raw[child:child + 4] == b'SNOD'
            nsym = struct.unpack_from('<H', raw, child + 6)[0]
            assert 1 <= nsym <= 8
            part = [f._heap_name(hp, struct.unpack_from('<Q', raw, child + 8 + 40 * k)[0]) for k in range(nsym)]
            assert f._heap_name(hp, key_hi) == part[-1]                   # key = largest name of the child
            names += part
        assert names == sorted(names, key=lambda s: s.encode()) and len(set(names)) == len(names)
        for n in names:
            ch = node[n]
            if ch.is_group:
                group(ch)
            else:
                dataset(ch)

    def dataset(node):
        header(node.addr)
        count[0] += 1
        lay = node.layout
        assert lay['class'] == 'chunked' and node.maxshape == (None,) and lay['elem'] == node.dtype.itemsize
        idx = f.chunk_index(lay)
        rows, c = node.shape[0], lay['chunk'][0]
        assert [o[0] for o, _, _, _ in idx] == list(range(0, len(idx) * c, c)) and len(idx) == -(-rows // c)
        for _, addr, nbytes, mask in idx:
            assert addr % 8 == 0 and nbytes == c * lay['elem'] and mask == 0 and addr + nbytes <= size
        if rows:
            levels(lay['btree'], c)

    def levels(root, c):
        """sibling links and the closing key of every level of a chunk B-tree"""
        addr = root
        while True:
            typ, level, used = struct.unpack_from('<BBH', raw, addr + 4)
            assert typ == 1
            a, prev = addr, h5lite.UNDEF
            last_key = -1
            while a != h5lite.UNDEF:
                t2, l2, u2 = struct.unpack_from('<BBH', raw, a + 4)
                left, right = struct.unpack_from('<QQ', raw, a + 8)
                assert (t2, l2) == (1, level) and left == prev and 1 <= u2 <= 64
                keys = [struct.unpack_from('<Q', raw, a + 24 + 32 * i + 8)[0] for i in range(u2 + 1)]
                assert keys == sorted(keys) and keys[0] > last_key
                last_key = keys[-2]
                prev, a = a, right
            if level == 0:
                return
            addr = struct.unpack_from('<Q', raw, addr + 24 + 24)[0]       # first child of the leftmost node

    group(f)
    return count[0]


@pytest.mark.parametrize('rows,chunk', [(0, 16), (5, 16), (16, 16), (1000, 7), (70000, 16), (300000, 64)])
def test_table_roundtrip_and_btree_depth(tmp_path, rows, chunk):
    dt = np.dtype([('index', '<i8'), ('type', 'S16'), ('energy', '<f8'), ('x', '<f4')])
    rng = np.random.default_rng(rows)
    data = np.zeros(rows, dt)
    data['index'] = np.arange(rows)
    data['type'] = np.array([b'Ruth', b'Moller', b'Gryz3s3p3d', b'Quinn'])[rng.integers(0, 4, rows)]
    data['energy'] = rng.uniform(1e4, 2e4, rows)
    data['x'] = rng.normal(size=rows)
    path = str(tmp_path / 't.h5')
    with h5lite.Writer(path) as w:
        t = w.create_table('/scats/table', dt, chunk_rows=chunk)
        for part in np.array_split(data, 13):                            # ragged appends across chunk boundaries
            w.append(t, part)
        w.set_attrs('/scats/table', {'NROWS': np.int64(rows), 'TITLE': '', 'CLASS': 'TABLE'})
    assert _validate(path) == 3
    f = h5lite.File(path)
    tab = f['scats/table']
    assert tab.shape == (rows,) and tab.dtype == dt
    assert np.array_equal(tab.read(), data)
    assert tab.attr('NROWS') == rows and tab.attr('TITLE') == '' and tab.attr('CLASS') == 'TABLE'
    nchunks = -(-rows // chunk)
    depth = 0 if nchunks == 0 else 1 if nchunks <= 64 else 2 if nchunks <= 4096 else 3
    if depth:
        assert open(path, 'rb').read()[tab.layout['btree'] + 5] == depth - 1     # level of the root node


def test_groups_attributes_and_many_members(tmp_path):
    path = str(tmp_path / 'g.h5')
    dt = np.dtype([('index', '<i8'), ('v', '<f8')])
    attrs = {'CLASS': 'GROUP', 'levels': 1, 'ratio': 0.25, 'raw': b'bytes', 'empty': '', 'uni': 'Å-units',
             'index_cols': [(0, 'index')], 'info': {1: {'type': 'Index', 'names': [None]}, 'index': {}},
             'none': None, 'np': np.float32(1.5), 'u8': np.uint8(7)}
    with h5lite.Writer(path) as w:
        for k in range(40):                                              # 5 symbol-table nodes in one group
            t = w.create_table('/many/t%02d' % k, dt, chunk_rows=4)
            w.append(t, np.array([(k, k * 0.5)], dt))
        w.create_group('/a/b/c')
        w.set_attrs('/a/b', attrs)
        w.set_attrs('/', {'PYTABLES_FORMAT_VERSION': '2.1'})
    assert _validate(path) == 1 + 1 + 40 + 3
    f = h5lite.File(path)
    assert f.keys() == ['a', 'many'] and f['many'].keys() == ['t%02d' % k for k in range(40)]
    assert f['many/t17'].read()['v'][0] == 8.5 and f['a/b'].keys() == ['c'] and f['a/b/c'].keys() == []
    g = f['a/b']
    for k, v in attrs.items():
        assert g.attr(k) == v, k
    assert g.attrs['index_cols'] == pickle.dumps([(0, 'index')], 0)       # PyTables: protocol-0 pickle in a string
    assert f.attr('PYTABLES_FORMAT_VERSION') == '2.1'
    with pytest.raises(KeyError):
        f['a/nope']


def test_store_writes_the_pandas_frame_table_contract(tmp_path):
    """What pandas.HDFStore.put(key, df, format='table', data_columns=True) leaves in a file (io/pytables.py:
    set_object_info, Table.set_attrs, IndexCol/DataCol.set_attr) and PyTables adds for a Table."""
    from test_store_cli import _fake_result, IP, NAMES
    from mcdiscrete_b200.store import ResultStore, read_store, STR_WIDTH
    res, rows = _fake_result()
    path = str(tmp_path / 'run.h5')
    with ResultStore(path, IP, fmt='hdf5', chunk_rows=4) as st:
        assert st.fmt in ('h5lite', 'hdf5')
        if st.fmt != 'h5lite':
            pytest.skip('PyTables is installed: pandas writes the file itself')
        st.append_result(res, NAMES)
        st.append_result(res, NAMES, double_record=True)
    _validate(path)
    f = h5lite.File(path)
    assert f.keys() == ['els', 'input', 'scats']
    assert f.attr('CLASS') == 'GROUP' and f.attr('PYTABLES_FORMAT_VERSION') == '2.1'
    for key, cols in (('els', ['outcome', 'final_E', 'dx', 'dy', 'dz', 'last_x', 'last_y', 'last_z']),
                      ('scats', ['type', 'energy', 'x', 'y', 'z'])):
        g, t = f[key], f[key + '/table']
        assert g.attr('pandas_type') == 'frame_table' and g.attr('table_type') == 'appendable_frame'
        assert g.attr('pandas_version') == '0.15.2' and g.attr('encoding') == 'UTF-8' and g.attr('nan_rep') == 'nan'
        assert g.attr('index_cols') == [(0, 'index')] and g.attr('levels') == 1
        assert g.attr('values_cols') == cols == g.attr('data_columns') and g.attr('non_index_axes') == [(1, cols)]
        assert g.attr('info') == {1: {'type': 'Index', 'names': [None]}, 'index': {}}
        assert t.attr('CLASS') == 'TABLE' and t.attr('VERSION') == '2.7' and t.attr('NROWS') == t.shape[0]
        assert t.dtype.names == tuple(['index'] + cols) and t.attr('index_kind') == 'integer'
        for i, n in enumerate(t.dtype.names):
            assert t.attr('FIELD_%d_NAME' % i) == n
        for c in cols:
            assert t.attr(c + '_kind') == [c] and t.attr(c + '_meta') is None
            assert t.attr(c + '_dtype') == ('bytes%d' % (8 * STR_WIDTH) if c in ('outcome', 'type') else 'float64')
        assert np.array_equal(t.read()['index'], np.arange(t.shape[0]))
    back = read_store(path)
    assert back['input']['material'] == 'Si' and back['input']['num_el'] == '5'
    assert list(back['els']['outcome']) == ['abs', 'bks', 'bks', 'long', 'abs'] + ['abs', 'bks', 'bks', 'abs', 'long', 'abs']
    assert back['els']['final_E'].dtype == np.float64 and len(back['scats']) == 2 * len(rows)
    assert list(back['scats']['type'][:3]) == ['Ruth', 'Moller', 'Gryz1s']
    assert list(back['scats']['x']) == 2 * [float(r[2]) for r in rows]


def test_h5lite_file_opens_with_pandas_where_pytables_exists(tmp_path):
    """The compatibility check this image cannot run (no PyTables / libhdf5 here): wherever `tables` is installed, a
    store written by the self-contained writer must open with pandas.read_hdf exactly like one written by pandas."""
    pytest.importorskip('tables')
    import pandas as pd
    from test_store_cli import _fake_result, IP, NAMES
    from mcdiscrete_b200.store import ResultStore, read_store
    res, rows = _fake_result()
    path = str(tmp_path / 'lite.h5')
    with ResultStore(path, IP, fmt='h5lite', chunk_rows=4) as st:
        st.append_result(res, NAMES)
        st.append_result(res, NAMES)
    mine = read_store(path, reader='h5lite')
    for key in ('els', 'scats'):
        df = pd.read_hdf(path, key)
        assert list(df.columns) == list(mine[key].columns) and len(df) == len(mine[key])
        for c in df.columns:
            assert list(df[c]) == list(mine[key][c]), (key, c)
    assert pd.read_hdf(path, 'input').iloc[0]['material'] == 'Si'
    sel = pd.read_hdf(path, 'scats', where='type == "Ruth"')          # data columns are queryable
    assert len(sel) == int((mine['scats']['type'] == 'Ruth').sum())


def test_small_tables_fit_exactly_and_describe(tmp_path):
    """A table smaller than one chunk is stored with a chunk of exactly its size (no 5 MB file for 5 rows), and
    `python -m mcdiscrete_b200.h5lite file` prints the tree."""
    import io
    dt = np.dtype([('index', '<i8'), ('v', '<f8')])
    path = str(tmp_path / 's.h5')
    with h5lite.Writer(path) as w:
        t = w.create_table('/els/table', dt, chunk_rows=1 << 16)
        w.append(t, np.array([(0, 1.0), (1, 2.0), (2, 3.0)], dt))
    assert os.path.getsize(path) < 8192              # fixed-size B-tree nodes dominate (2 KB + 2 x 0.5 KB)
    f = h5lite.File(path)
    assert f['els/table'].layout['chunk'] == (3,) and list(f['els/table'].read()['v']) == [1.0, 2.0, 3.0]
    _validate(path)
    buf = io.StringIO()
    h5lite.describe(path, buf)
    assert 'els/' in buf.getvalue() and 'shape (3,), chunked' in buf.getvalue()


def test_roundtrip_property_random_tables(tmp_path):
    """Property test (hypothesis): any mix of int / float / fixed-string columns, any row count and chunk size, any
    split into appends -> the file validates and reads back bit for bit."""
    hyp = pytest.importorskip('hypothesis')
    from hypothesis import given, settings, strategies as st_

    kinds = st_.sampled_from(['<i8', '<i4', '<u2', '<f8', '<f4', 'S1', 'S7', 'S16'])
    counter = [0]

    @settings(max_examples=40, deadline=None)
    @given(cols=st_.lists(kinds, min_size=1, max_size=6), rows=st_.integers(0, 700), chunk=st_.integers(1, 97),
           parts=st_.integers(1, 9), seed=st_.integers(0, 2 ** 31))
    def check(cols, rows, chunk, parts, seed):
        rng = np.random.default_rng(seed)
        dt = np.dtype([('index', '<i8')] + [('c%d' % i, k) for i, k in enumerate(cols)])
        data = np.zeros(rows, dt)
        data['index'] = np.arange(rows)
        for i, k in enumerate(cols):
            sub = np.dtype(k)
            if sub.kind == 'S':
                raw = rng.integers(97, 123, (rows, sub.itemsize), dtype=np.uint8)
                raw[rng.random((rows, sub.itemsize)) < 0.2] = 0          # embedded / trailing NULs
                data['c%d' % i] = np.frombuffer(raw.tobytes(), dtype=sub, count=rows) if rows else np.zeros(0, sub)
            elif sub.kind == 'f':
                data['c%d' % i] = rng.normal(size=rows).astype(sub)
            else:
                info = np.iinfo(sub)
                data['c%d' % i] = rng.integers(info.min, info.max, rows, dtype=sub, endpoint=True)
        counter[0] += 1
        path = str(tmp_path / ('p%d.h5' % counter[0]))
        with h5lite.Writer(path) as w:
            t = w.create_table('/g/table', dt, chunk_rows=chunk)
            for part in np.array_split(data, parts):
                w.append(t, part)
            w.set_attrs('/g/table', {'NROWS': np.int64(rows)})
        _validate(path)
        back = h5lite.File(path)['g/table']
        assert back.dtype == dt and back.shape == (rows,)
        assert back.read().tobytes() == data.tobytes()
        os.remove(path)

    check()
