"""Self-contained subset of the HDF5 file format (no libhdf5, no PyTables, no h5py): enough to WRITE the result
store of the reference -- `pd.HDFStore.put(key, df, format='table', data_columns=True, append=True)`
(MC/doScatter.py:63-67, 112-117) -- and to read such files back.  SURVEY.md 8(f) N1(b).

Written from the published HDF5 File Format Specification (version 1.1 structures, the ones every libhdf5 since 1.6
reads): superblock version 0, version-1 object headers, "old-style" groups (symbol-table message -> version-1
B-tree + local heap + symbol-table nodes), chunked datasets indexed by a version-1 chunk B-tree with an unlimited
first dimension (what PyTables creates for a Table), compound datatypes (version 1) of little-endian integers,
IEEE floats and fixed-length strings, version-1 attribute messages.  On top of that the PyTables / pandas attribute
conventions (CLASS, VERSION, TITLE, FIELD_n_NAME, NROWS; pandas_type, table_type, index_cols, values_cols,
non_index_axes, data_columns, <col>_kind, <col>_dtype, <col>_meta ...; non-scalar values as protocol-0 pickles in
fixed-length strings, as PyTables stores them).

This image has no HDF5 library, so what is verified here is: (1) the reader below parses a file written by the
real libhdf5 (MATLAB 7.3 test file shipped inside scipy, tests/test_h5lite.py) -- it is an independent judge of the
format; (2) everything the writer produces is re-parsed by that reader and compared value for value.  Opening the
files with pandas / PyTables could not be tried here and is documented as unverified.

Writer layout of a file:  superblock | chunks of every table as they are appended (streaming, one write each) |
at close: chunk B-trees, object headers, group B-trees / heaps / symbol nodes; then the superblock is patched
(root object header, end-of-file address).
"""
import pickle
import struct
import zlib

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xffffffffffffffff
GROUP_LEAF_K = 4           # symbol-table nodes hold up to 2K = 8 entries
GROUP_INTERNAL_K = 16      # group B-tree nodes hold up to 2K = 32 children
CHUNK_K = 32               # chunk B-tree nodes hold up to 2K = 64 children (library default; superblock v0 has no field)


def _pad8(b):
    return b + b'\x00' * (-len(b) % 8)


# =====================================================================================================
# datatype messages <-> numpy dtypes
# =====================================================================================================
def encode_dtype(dt, utf8=False):
    """numpy dtype -> HDF5 datatype message (version 1 encodings)."""
    dt = np.dtype(dt)
    if dt.names:
        out = struct.pack('<BBBBI', 0x16, len(dt.names) & 0xff, len(dt.names) >> 8, 0, dt.itemsize)
        for name in dt.names:
            sub, off = dt.fields[name][:2]
            nm = name.encode('utf-8') + b'\x00'
            out += nm + b'\x00' * (-len(nm) % 8)
            out += struct.pack('<IB3xI4x4I', off, 0, 0, 0, 0, 0, 0)
            out += encode_dtype(sub)
        return out
    if dt.kind in 'iu':
        return struct.pack('<BBBBIHH', 0x10, 0x08 if dt.kind == 'i' else 0x00, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == 'f':
        if dt.itemsize == 8:
            return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
        if dt.itemsize == 4:
            return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x10 if utf8 else 0x00, 0, 0, max(dt.itemsize, 1))
    raise TypeError('h5lite cannot encode dtype %r' % (dt,))


def decode_dtype(b, p=0):
    """HDF5 datatype message at b[p:] -> (numpy dtype or None, bytes consumed, info dict)."""
    if not isinstance(b, (bytes, bytearray)):
        b = bytes(b)
    cv, b0, b1, b2, size = struct.unpack_from('<BBBBI', b, p)
    cls, ver = cv & 0x0f, cv >> 4
    q = p + 8
    info = {'class': cls, 'version': ver, 'size': size}
    if cls == 0:                                                      # fixed point
        q += 4
        return np.dtype('%s%s%d' % ('>' if b0 & 1 else '<', 'i' if b0 & 8 else 'u', size)), q - p, info
    if cls == 1:                                                      # floating point
        q += 12
        return np.dtype('%sf%d' % ('>' if b0 & 1 else '<', size)), q - p, info
    if cls == 3:                                                      # string
        info['cset'] = 'utf-8' if (b0 >> 4) == 1 else 'ascii'
        return np.dtype('S%d' % size), q - p, info
    if cls == 4:                                                      # bit field
        q += 4
        return np.dtype('u%d' % size), q - p, info
    if cls == 5:                                                      # opaque
        q += (b0 + 7) // 8 * 8
        return np.dtype('V%d' % size), q - p, info
    if cls == 7:                                                      # reference
        return np.dtype('<u%d' % size) if size in (1, 2, 4, 8) else np.dtype('V%d' % size), q - p, info
    if cls == 6:                                                      # compound
        nmem = b0 | (b1 << 8)
        names, fmts, offs = [], [], []
        for _ in range(nmem):
            e = b.index(b'\x00', q)
            name = bytes(b[q:e]).decode('utf-8')
            q = (q + (e - q + 8) // 8 * 8) if ver < 3 else e + 1
            if ver < 3:
                off, ndim = struct.unpack_from('<IB', b, q)
                dims = struct.unpack_from('<4I', b, q + 16)
                q += 12 if ver == 2 else 32
                if ver == 2:
                    ndim = 0
            else:
                nb = 1 if size < 256 else 2 if size < 65536 else 3 if size < (1 << 24) else 4
                off = int.from_bytes(bytes(b[q:q + nb]), 'little')
                q += nb
                ndim = 0
            sub, used, _ = decode_dtype(b, q)
            q += used
            if ver == 1 and ndim:
                sub = np.dtype((sub, tuple(dims[:ndim])))
            names.append(name)
            fmts.append(sub if sub is not None else np.dtype('V1'))
            offs.append(off)
        return np.dtype({'names': names, 'formats': fmts, 'offsets': offs, 'itemsize': size}), q - p, info
    if cls == 8:                                                      # enumeration: base type + names + values
        nmem = b0 | (b1 << 8)
        base, used, _ = decode_dtype(b, q)
        q += used
        for _ in range(nmem):
            e = b.index(b'\x00', q)
            q = (q + (e - q + 8) // 8 * 8) if ver < 3 else e + 1
        q += nmem * base.itemsize
        return base, q - p, info
    if cls == 9:                                                      # variable length: not materialised
        _, used, _ = decode_dtype(b, q)
        return None, q + used - p, info
    if cls == 10:                                                     # array
        ndim = b[q]
        q += 4 if ver < 3 else 1
        dims = struct.unpack_from('<%dI' % ndim, b, q)
        q += 4 * ndim * (2 if ver < 3 else 1)
        base, used, _ = decode_dtype(b, q)
        return np.dtype((base, tuple(dims))), q + used - p, info
    raise ValueError('datatype class %d is not supported' % cls)


# =====================================================================================================
# reader
# =====================================================================================================
class Node:
    """A group or a dataset of an opened file."""

    def __init__(self, f, addr, name):
        self.file, self.addr, self.name = f, addr, name
        self.msgs = f._object_header(addr)
        self._attrs = None

    # ---- kind ------------------------------------------------------------------------------------
    @property
    def is_group(self):
        return any(t == 0x11 for t, _, _ in self.msgs)

    def _msg(self, typ):
        for t, _, d in self.msgs:
            if t == typ:
                return d
        return None

    # ---- attributes ------------------------------------------------------------------------------
    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = {}
            for t, _, d in self.msgs:
                if t == 0x0c:
                    k, v = self.file._attribute(d)
                    self._attrs[k] = v
        return self._attrs

    def attr(self, name, unpickle=True):
        """Attribute value the way PyTables hands it out: byte strings that look like pickles are loaded."""
        v = self.attrs[name]
        if unpickle and isinstance(v, bytes) and v.endswith(b'.') and v not in (b'0', b'0.'):
            try:
                return pickle.loads(v)
            except Exception:
                return v
        return v

    # ---- group -----------------------------------------------------------------------------------
    def keys(self):
        return list(self._links())

    def _links(self):
        d = self._msg(0x11)
        if d is None:
            raise TypeError('%s is not a group' % self.name)
        btree, heap = struct.unpack_from('<QQ', d, 0)
        return self.file._group_entries(btree, heap)

    def __contains__(self, k):
        return k in self._links()

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split('/') if p]:
            links = node._links()
            if part not in links:
                raise KeyError('%s has no member %r' % (node.name, part))
            node = Node(self.file, links[part], node.name.rstrip('/') + '/' + part)
        return node

    # ---- dataset ---------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.file._dataspace(self._msg(0x01))[0]

    @property
    def maxshape(self):
        return self.file._dataspace(self._msg(0x01))[1]

    @property
    def dtype(self):
        return decode_dtype(self._msg(0x03))[0]

    @property
    def layout(self):
        return self.file._layout(self._msg(0x08))

    def read(self):
        """The whole dataset as a numpy array."""
        f = self.file
        shape = self.shape
        dt = self.dtype
        if dt is None:
            raise TypeError('variable-length data are not supported')
        n = int(np.prod(shape)) if shape else 1
        lay = self.layout
        if lay['class'] == 'compact':
            raw = lay['data']
        elif lay['class'] == 'contiguous':
            raw = b'' if lay['address'] == UNDEF else f._bytes(lay['address'], n * dt.itemsize)
        else:
            raw = f._read_chunked(lay, shape, dt, self._msg(0x0b))
        return np.frombuffer(bytes(raw), dtype=dt, count=n).reshape(shape)


class File(Node):
    """Read-only view of an HDF5 file (the subset described in the module docstring)."""

    def __init__(self, path):
        self.path = path
        self.buf = np.memmap(path, dtype=np.uint8, mode='r') if _size(path) else np.zeros(0, np.uint8)
        off = 0
        while True:
            if off + 8 > len(self.buf):
                raise ValueError('%s: no HDF5 signature' % path)
            if bytes(self.buf[off:off + 8]) == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
        b = self.buf
        ver = int(b[off + 8])
        if ver > 1:
            raise ValueError('superblock version %d is not supported (only 0 and 1)' % ver)
        self.superblock_version = ver
        self.size_of_offsets, self.size_of_lengths = int(b[off + 13]), int(b[off + 14])
        if (self.size_of_offsets, self.size_of_lengths) != (8, 8):
            raise ValueError('only 8-byte offsets and lengths are supported')
        self.group_leaf_k, self.group_internal_k = struct.unpack_from('<HH', b, off + 16)
        p = off + 24 + (4 if ver == 1 else 0)
        self.base, self.free_addr, self.eof, self.driver_addr = struct.unpack_from('<QQQQ', b, p)
        p += 32
        _, ohdr, cache, _, bt, hp = struct.unpack_from('<QQIIQQ', b, p)
        self.root_cache = (cache, bt, hp)
        self.file = self
        Node.__init__(self, self, ohdr, '/')

    def close(self):
        self.buf = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False

    # ---- raw access ------------------------------------------------------------------------------
    def _bytes(self, addr, n):
        a = self.base + addr
        if a + n > len(self.buf):
            raise ValueError('address range %d+%d beyond the end of the file' % (a, n))
        return self.buf[a:a + n]

    def _object_header(self, addr):
        b = self._bytes(addr, 16)
        if bytes(b[:4]) == b'OHDR':
            raise ValueError('version-2 object headers are not supported')
        ver, _, nmsg, _, size = struct.unpack_from('<BBHII', b, 0)
        if ver != 1:
            raise ValueError('object header version %d at %d' % (ver, addr))
        msgs = []
        blocks = [(addr + 16, size)]
        while blocks and len(msgs) < nmsg:
            a, n = blocks.pop(0)
            blk = self._bytes(a, n)
            p = 0
            while p + 8 <= n and len(msgs) < nmsg:
                t, sz, fl = struct.unpack_from('<HHB', blk, p)
                d = blk[p + 8:p + 8 + sz]
                p += 8 + sz
                if t == 0x10:                                          # continuation
                    ca, cn = struct.unpack_from('<QQ', d, 0)
                    blocks.append((ca, cn))
                msgs.append((t, fl, d))
        return msgs

    def _heap_name(self, heap, off):
        h = self._bytes(heap, 32)
        if bytes(h[:4]) != b'HEAP':
            raise ValueError('bad local heap signature')
        seg_size, _, seg = struct.unpack_from('<QQQ', h, 8)
        d = self._bytes(seg, seg_size)
        e = off
        while d[e]:
            e += 1
        return bytes(d[off:e]).decode('utf-8')

    def _group_entries(self, btree, heap):
        out = {}

        def walk(addr):
            b = self._bytes(addr, 24)
            if bytes(b[:4]) != b'TREE':
                raise ValueError('bad B-tree signature at %d' % addr)
            typ, level, used = struct.unpack_from('<BBH', b, 4)
            if typ != 0:
                raise ValueError('not a group B-tree')
            body = self._bytes(addr + 24, (2 * used + 1) * 8)
            for i in range(used):
                child = struct.unpack_from('<Q', body, 8 + 16 * i)[0]
                if level:
                    walk(child)
                else:
                    s = self._bytes(child, 8)
                    if bytes(s[:4]) != b'SNOD':
                        raise ValueError('bad symbol-table node signature')
                    nsym = struct.unpack_from('<H', s, 6)[0]
                    ents = self._bytes(child + 8, 40 * nsym)
                    for k in range(nsym):
                        noff, ohdr = struct.unpack_from('<QQ', ents, 40 * k)
                        out[self._heap_name(heap, noff)] = ohdr
        if btree != UNDEF:
            walk(btree)
        return out

    @staticmethod
    def _dataspace(d):
        ver, rank, flags = int(d[0]), int(d[1]), int(d[2])
        p = 8 if ver == 1 else 4
        dims = struct.unpack_from('<%dQ' % rank, d, p)
        mx = struct.unpack_from('<%dQ' % rank, d, p + 8 * rank) if flags & 1 else dims
        return tuple(int(x) for x in dims), tuple(None if x == UNDEF else int(x) for x in mx)

    @staticmethod
    def _layout(d):
        ver = int(d[0])
        if ver == 3:
            cls = int(d[1])
            if cls == 0:
                n = struct.unpack_from('<H', d, 2)[0]
                return {'class': 'compact', 'data': d[4:4 + n]}
            if cls == 1:
                a, n = struct.unpack_from('<QQ', d, 2)
                return {'class': 'contiguous', 'address': a, 'size': n}
            if cls == 2:
                nd = int(d[2])
                a = struct.unpack_from('<Q', d, 3)[0]
                dims = struct.unpack_from('<%dI' % nd, d, 11)
                return {'class': 'chunked', 'btree': a, 'chunk': tuple(dims[:-1]), 'elem': dims[-1]}
        elif ver in (1, 2):
            nd, cls = int(d[1]), int(d[2])
            p = 8
            a = UNDEF
            if cls != 0:
                a = struct.unpack_from('<Q', d, p)[0]
                p += 8
            dims = struct.unpack_from('<%dI' % nd, d, p)
            p += 4 * nd
            if cls == 0:
                n = struct.unpack_from('<I', d, p)[0]
                return {'class': 'compact', 'data': d[p + 4:p + 4 + n]}
            if cls == 1:
                return {'class': 'contiguous', 'address': a, 'size': None}
            return {'class': 'chunked', 'btree': a, 'chunk': tuple(dims[:-1]), 'elem': dims[-1]}
        raise ValueError('data layout version %d / class is not supported' % ver)

    def chunk_index(self, lay):
        """[(offset tuple, address, stored bytes, filter mask)] of a chunked dataset, in B-tree order."""
        nd = len(lay['chunk']) + 1
        ksz = 8 + 8 * nd
        out = []

        def walk(addr):
            b = self._bytes(addr, 24)
            if bytes(b[:4]) != b'TREE':
                raise ValueError('bad chunk B-tree signature at %d' % addr)
            typ, level, used = struct.unpack_from('<BBH', b, 4)
            if typ != 1:
                raise ValueError('not a chunk B-tree')
            body = self._bytes(addr + 24, used * (ksz + 8) + ksz)
            for i in range(used):
                p = i * (ksz + 8)
                nbytes, mask = struct.unpack_from('<II', body, p)
                offs = struct.unpack_from('<%dQ' % nd, body, p + 8)
                child = struct.unpack_from('<Q', body, p + ksz)[0]
                if level:
                    walk(child)
                else:
                    out.append((tuple(int(x) for x in offs[:-1]), child, nbytes, mask))
        if lay['btree'] != UNDEF:
            walk(lay['btree'])
        return out

    @staticmethod
    def _filters(d):
        if d is None:
            return []
        ver, nf = int(d[0]), int(d[1])
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(nf):
            fid = struct.unpack_from('<H', d, p)[0]
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from('<H', d, p + 2)[0]
                p += 4
            else:
                nlen = 0
                p += 2
            _, ncd = struct.unpack_from('<HH', d, p)
            p += 4 + (((nlen + 7) // 8 * 8) if ver == 1 else nlen)
            cd = struct.unpack_from('<%dI' % ncd, d, p)
            p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
            out.append((fid, cd))
        return out

    def _read_chunked(self, lay, shape, dt, filt_msg):
        filters = self._filters(filt_msg)
        out = np.zeros(shape, dtype=dt)
        cshape = lay['chunk']
        for offs, addr, nbytes, mask in self.chunk_index(lay):
            raw = bytes(self._bytes(addr, nbytes))
            for k, (fid, cd) in reversed(list(enumerate(filters))):
                if mask & (1 << k):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0]
                    a = np.frombuffer(raw, np.uint8)
                    m = len(a) // es
                    raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                elif fid == 3:                                         # fletcher32: strip the checksum
                    raw = raw[:-4]
                else:
                    raise ValueError('filter %d is not supported' % fid)
            chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cshape))).reshape(cshape)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out

    def _attribute(self, d):
        ver = int(d[0])
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from('<HHH', d, 2)
            p = 8
            name = bytes(d[p:p + nsz]).split(b'\x00')[0].decode('utf-8')
            p += (nsz + 7) // 8 * 8
            tp, sp = p, p + (tsz + 7) // 8 * 8
            p = sp + (ssz + 7) // 8 * 8
        elif ver in (2, 3):
            nsz, tsz, ssz = struct.unpack_from('<HHH', d, 2)
            p = 8 + (1 if ver == 3 else 0)
            name = bytes(d[p:p + nsz]).split(b'\x00')[0].decode('utf-8')
            tp = p + nsz
            sp = tp + tsz
            p = sp + ssz
        else:
            raise ValueError('attribute message version %d' % ver)
        dt, _, info = decode_dtype(d, tp)
        sver, rank = int(d[sp]), int(d[sp + 1])
        dims = struct.unpack_from('<%dQ' % rank, d, sp + (8 if sver == 1 else 4)) if rank else ()
        if sver == 2 and int(d[sp + 3]) == 2:                          # null dataspace
            return name, None
        if dt is None:
            return name, None
        n = int(np.prod(dims)) if rank else 1
        arr = np.frombuffer(bytes(d[p:p + n * dt.itemsize]), dtype=dt, count=n)
        if not rank:
            v = arr[0]
            if dt.kind == 'S':
                v = bytes(v)
                return name, (v.decode('utf-8') if info.get('cset') == 'utf-8' else v)
            return name, v.item() if dt.names is None else v
        return name, arr.reshape(dims)


def _size(path):
    import os
    return os.path.getsize(path)


# =====================================================================================================
# writer
# =====================================================================================================
def _attr_value(v):
    """python value -> (datatype message, dataspace message, data bytes) following PyTables' rules: str ->
    UTF-8 string scalar, bytes -> ASCII string scalar, bool/int/float -> numpy scalar, numpy scalars as they
    are, anything else -> protocol-0 pickle in an ASCII string."""
    scalar_space = struct.pack('<BBB5x', 1, 0, 0)
    if isinstance(v, str):
        b = v.encode('utf-8') or b'\x00'
        return encode_dtype(np.dtype('S%d' % len(b)), utf8=True), scalar_space, b
    if isinstance(v, bytes):
        b = v or b'\x00'
        return encode_dtype(np.dtype('S%d' % len(b))), scalar_space, b
    if isinstance(v, (bool, np.bool_)):
        v = np.uint8(v)          # PyTables stores numpy.bool_ as an HDF5 enum; a plain byte keeps this subset small
    if isinstance(v, int):
        v = np.int64(v)
    if isinstance(v, float):
        v = np.float64(v)
    if isinstance(v, np.generic) and v.dtype.kind in 'iuf':
        return encode_dtype(v.dtype), scalar_space, v.tobytes()
    return _attr_value(pickle.dumps(v, 0))


def _attr_msg(name, v):
    t, s, data = _attr_value(v)
    nm = name.encode('utf-8') + b'\x00'
    body = struct.pack('<BxHHH', 1, len(nm), len(t), len(s)) + _pad8(nm) + _pad8(t) + _pad8(s) + data
    return (0x0c, body)


def _object_header(msgs):
    body = b''
    for typ, data in msgs:
        data = _pad8(data)
        if len(data) > 0xffff:
            raise ValueError('header message of %d bytes is too large' % len(data))
        body += struct.pack('<HHB3x', typ, len(data), 0) + data
    return struct.pack('<BxHII4x', 1, len(msgs), 1, len(body)) + body


class _Table:
    def __init__(self, dtype, chunk_rows):
        self.dtype = np.dtype(dtype)
        self.chunk_rows = int(chunk_rows)
        self.rows = 0
        self.chunks = []             # file address of chunk k (rows k*chunk_rows ...)
        self.pending = []            # arrays not yet forming a full chunk
        self.n_pending = 0
        self.attrs = {}


class Writer:
    """Write an HDF5 file of groups and appendable tables (1-D chunked compound datasets, unlimited length).

        w = Writer(path)
        t = w.create_table('/els/table', dtype, chunk_rows=65536)
        w.append(t, rows); ...          # streaming: every full chunk goes to the file at once
        w.set_attrs('/els', {...}); w.set_attrs('/els/table', {...})
        w.close()
    """

    def __init__(self, path):
        self.path = path
        self.f = open(path, 'wb')
        self.f.write(b'\x00' * 96)                                     # superblock, patched at close
        self.pos = 96
        self.tree = {'children': {}, 'attrs': {}}                       # nested groups; leaves are _Table
        self.closed = False

    # ---- file space ------------------------------------------------------------------------------
    def _write(self, b):
        """Append 8-byte aligned; returns the address."""
        pad = -self.pos % 8
        if pad:
            self.f.write(b'\x00' * pad)
            self.pos += pad
        a = self.pos
        self.f.write(b)
        self.pos += len(b)
        return a

    # ---- tree ------------------------------------------------------------------------------------
    def _group(self, parts, create=True):
        g = self.tree
        for p in parts:
            ch = g['children']
            if p not in ch:
                if not create:
                    raise KeyError(p)
                ch[p] = {'children': {}, 'attrs': {}}
            g = ch[p]
            if isinstance(g, _Table):
                raise ValueError('%s is a table' % p)
        return g

    def create_group(self, path):
        return self._group([p for p in path.split('/') if p])

    def create_table(self, path, dtype, chunk_rows=65536):
        parts = [p for p in path.split('/') if p]
        g = self._group(parts[:-1])
        if parts[-1] in g['children']:
            raise ValueError('%s exists' % path)
        t = g['children'][parts[-1]] = _Table(dtype, chunk_rows)
        return t

    def set_attrs(self, path, attrs):
        parts = [p for p in path.split('/') if p]
        node = self.tree
        for p in parts:
            node = node['children'][p]
        (node.attrs if isinstance(node, _Table) else node['attrs']).update(attrs)

    # ---- rows ------------------------------------------------------------------------------------
    def append(self, t, rows):
        rows = np.ascontiguousarray(rows, dtype=t.dtype)
        if not len(rows):
            return
        t.rows += len(rows)
        c = t.chunk_rows
        if t.n_pending:
            take = min(c - t.n_pending, len(rows))
            t.pending.append(rows[:take])
            t.n_pending += take
            rows = rows[take:]
            if t.n_pending == c:
                t.chunks.append(self._write(np.concatenate(t.pending).tobytes()))
                t.pending, t.n_pending = [], 0
        full = len(rows) // c
        for k in range(full):
            t.chunks.append(self._write(rows[k * c:(k + 1) * c].tobytes()))
        rest = rows[full * c:]
        if len(rest):
            t.pending.append(rest.copy())
            t.n_pending += len(rest)

    # ---- close: metadata -------------------------------------------------------------------------
    def _chunk_btree(self, t):
        """Version-1 B-tree (node type 1) over the chunks; returns the root address (UNDEF without chunks)."""
        if not t.chunks:
            return UNDEF
        c, es = t.chunk_rows, t.dtype.itemsize
        nbytes = c * es
        ksz = 8 + 16
        node_size = 24 + 2 * CHUNK_K * 8 + (2 * CHUNK_K + 1) * ksz
        level = 0
        entries = [(k * c, a) for k, a in enumerate(t.chunks)]        # (first row, child address)
        end = len(t.chunks) * c
        while True:
            groups = [entries[i:i + 2 * CHUNK_K] for i in range(0, len(entries), 2 * CHUNK_K)]
            base = self._write(b'')                                    # aligned position of the first node
            addrs = [base + i * ((node_size + 7) // 8 * 8) for i in range(len(groups))]
            nxt = []
            for i, g in enumerate(groups):
                last = groups[i + 1][0][0] if i + 1 < len(groups) else end
                b = b'TREE' + struct.pack('<BBHQQ', 1, level, len(g), addrs[i - 1] if i else UNDEF,
                                          addrs[i + 1] if i + 1 < len(groups) else UNDEF)
                for row, child in g:
                    b += struct.pack('<IIQQQ', nbytes, 0, row, 0, child)
                b += struct.pack('<IIQQ', 0, 0, last, 0)
                b += b'\x00' * (node_size - len(b))
                a = self._write(b)
                assert a == addrs[i]
                nxt.append((g[0][0], a))
            if len(nxt) == 1:
                return nxt[0][1]
            entries, level = nxt, level + 1

    def _table_header(self, t):
        if t.n_pending and not t.chunks:
            t.chunk_rows = t.n_pending                                 # a table smaller than one chunk: exact fit
        if t.n_pending:                                                # last, partial chunk: stored at full size
            buf = np.zeros(t.chunk_rows, t.dtype)
            buf[:t.n_pending] = np.concatenate(t.pending)
            t.chunks.append(self._write(buf.tobytes()))
            t.pending, t.n_pending = [], 0
        bt = self._chunk_btree(t)
        space = struct.pack('<BBB5xQQ', 1, 1, 1, t.rows, UNDEF)
        fill = struct.pack('<BBBBI', 2, 3, 2, 1, 0)                    # v2: incremental allocation, fill if set, default value
        layout = struct.pack('<BBBQII', 3, 2, 2, bt, t.chunk_rows, t.dtype.itemsize)
        msgs = [(0x01, space), (0x03, encode_dtype(t.dtype)), (0x05, fill), (0x08, layout)]
        msgs += [_attr_msg(k, v) for k, v in t.attrs.items()]
        return self._write(_object_header(msgs))

    def _group_header(self, g):
        """Writes children first, then heap / symbol nodes / B-tree / header; returns (header, btree, heap)."""
        links = {}
        for name, ch in g['children'].items():
            if isinstance(ch, _Table):
                links[name] = (self._table_header(ch), None)
            else:
                links[name] = self._group_header(ch)
        names = sorted(links, key=lambda s: s.encode('utf-8'))
        # local heap: offset 0 = empty string; names 8-byte padded; one free block at the end
        seg = bytearray(8)
        off = {}
        for n in names:
            off[n] = len(seg)
            seg += _pad8(n.encode('utf-8') + b'\x00')
        free = len(seg)
        seg += struct.pack('<QQ', 1, 16)                               # free block: next = 1 (none), size 16
        seg_addr = self._write(bytes(seg))
        heap = self._write(b'HEAP' + struct.pack('<B3xQQQ', 0, len(seg), free, seg_addr))
        # symbol-table nodes of up to 2K entries, in name order
        per = 2 * GROUP_LEAF_K
        snods = []
        for i in range(0, len(names), per):
            part = names[i:i + per]
            b = b'SNOD' + struct.pack('<BxH', 1, len(part))
            for n in part:
                hdr, sub = links[n][0], links[n]
                if sub[1] is None:
                    b += struct.pack('<QQII16x', off[n], hdr, 0, 0)
                else:
                    b += struct.pack('<QQIIQQ', off[n], hdr, 1, 0, sub[1], sub[2])
            b += b'\x00' * (8 + per * 40 - len(b))
            snods.append((off[part[-1]], self._write(b)))
        if len(snods) > 2 * GROUP_INTERNAL_K:
            raise ValueError('more than %d members in one group are not supported' % (per * 2 * GROUP_INTERNAL_K))
        b = b'TREE' + struct.pack('<BBHQQ', 0, 0, len(snods), UNDEF, UNDEF) + struct.pack('<Q', 0)
        for last_name_off, a in snods:
            b += struct.pack('<QQ', a, last_name_off)
        b += b'\x00' * (24 + 2 * GROUP_INTERNAL_K * 8 + (2 * GROUP_INTERNAL_K + 1) * 8 - len(b))
        btree = self._write(b)
        msgs = [(0x11, struct.pack('<QQ', btree, heap))] + [_attr_msg(k, v) for k, v in g['attrs'].items()]
        return self._write(_object_header(msgs)), btree, heap

    def close(self):
        if self.closed:
            return
        root, btree, heap = self._group_header(self.tree)
        eof = self._write(b'')
        sb = SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
        sb += struct.pack('<QQIIQQ', 0, root, 1, 0, btree, heap)
        assert len(sb) == 96
        self.f.seek(0)
        self.f.write(sb)
        self.f.close()
        self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


def describe(path, out=None):
    """Print the tree of an HDF5 file this module can read (groups, datasets with shape / dtype / layout, number of
    attributes): `python -m mcdiscrete_b200.h5lite run.h5`."""
    import sys
    out = out or sys.stdout
    with File(path) as f:
        out.write('%s: superblock v%d, base address %d, end of file %d\n' % (path, f.superblock_version, f.base, f.eof))

        def walk(node, depth):
            for name in node.keys():
                ch = node[name]
                pad = '  ' * depth
                if ch.is_group:
                    out.write('%s%s/  (%d attributes)\n' % (pad, name, len(ch.attrs)))
                    walk(ch, depth + 1)
                else:
                    lay = ch.layout
                    extra = ', chunks of %s' % (lay['chunk'],) if lay['class'] == 'chunked' else ''
                    dt = ch.dtype
                    fields = ', '.join('%s:%s' % (n, dt.fields[n][0].str) for n in dt.names) if dt is not None and dt.names else str(dt)
                    out.write('%s%s  shape %s, %s%s, [%s]  (%d attributes)\n' % (pad, name, ch.shape, lay['class'], extra,
                                                                               fields, len(ch.attrs)))
        walk(f, 1)


if __name__ == '__main__':
    import sys
    for p in sys.argv[1:]:
        describe(p)
