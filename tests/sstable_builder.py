"""Independent pure-Python writer of big-format `oa` SSTables from logical rows (test helper; small inputs only).
Layouts follow SURVEY Appendix A: SortedTablePartitionWriter.java:97-166, UnfilteredSerializer.java:151-305, Cell.java:268-305,
ClusteringPrefix.java:455-477, BigFormatPartitionWriter.java:128-251, RowIndexEntry.java:625-642, IndexInfo.java:107-117."""
import struct
import oracle_lib as O
from cassandra_b200.io.sstable import SSTable, MARSHAL, type_class, column_class, is_complex, TIMESTAMP_EPOCH, DELETION_TIME_EPOCH, NO_DELETION_TIME
from cassandra_b200.io.compress import CompressionMetadata

LIVE = None
K_EXCL_END, K_INCL_START, K_EXCL_END_INCL_START, K_STATIC, K_CLUSTERING, K_INCL_END_EXCL_START, K_INCL_END, K_EXCL_START = range(8)
NO_TS = -(1 << 63)

def vint(v): return O.vint(v & ((1 << 64) - 1))
def i32s(v):            # writeUnsignedVInt32(int): sign-extended int
    v &= 0xFFFFFFFF
    if v >= 1 << 31: v -= 1 << 32
    return vint(v)

class Cell:
    """path: the cell path of a cell of a multi-cell column (map key / set element / list timeuuid), None for simple columns"""
    def __init__(self, col, ts, value=b"", ttl=0, ldt=NO_DELETION_TIME, path=None):
        self.col = col; self.ts = ts; self.value = value; self.ttl = ttl; self.ldt = ldt; self.path = path
    @staticmethod
    def tombstone(col, ts, ldt, path=None): return Cell(col, ts, b"", 0, ldt, path)
class Row:
    """complex_deletions: {column index: (markedForDeleteAt, localDeletionTime)} of multi-cell columns. The cells of a multi-cell column must be
    given in path order (the writer does not know the path comparator)."""
    def __init__(self, ck, cells=(), ts=NO_TS, ttl=0, ldt=NO_DELETION_TIME, deletion=LIVE, complex_deletions=None):
        self.ck = tuple(ck); self.cells = list(cells); self.ts = ts; self.ttl = ttl; self.ldt = ldt; self.deletion = deletion
        self.complex_deletions = dict(complex_deletions or {})
class Marker:
    """kind: bound/boundary kind ordinal; values: clustering prefix; close/open: (markedForDeleteAt, localDeletionTime) or None"""
    def __init__(self, kind, values, close=None, open=None):
        self.kind = kind; self.ck = tuple(values); self.close = close; self.open = open
class Partition:
    """static: Row with ck == () holding cells of the static columns (col = index in Schema.static_columns), or None"""
    def __init__(self, key, unfiltereds=(), deletion=LIVE, static=None):
        self.key = key; self.unfiltereds = list(unfiltereds); self.deletion = deletion; self.static = static

class Schema:
    def __init__(self, clustering_types, columns, static_columns=()):
        self.clustering_types = [t if "." in t else MARSHAL + t for t in clustering_types]
        self.columns = sorted([(n if isinstance(n, bytes) else n.encode(), t if "." in t else MARSHAL + t) for n, t in columns], key=lambda nt: (is_complex(nt[1]), nt[0]))
        self.static_columns = sorted([(n if isinstance(n, bytes) else n.encode(), t if "." in t else MARSHAL + t) for n, t in static_columns])
        self.cfixed = [type_class(t)[1] for t in self.clustering_types]
        self.vfixed = [column_class(t)[1] & 0xFFFF for _, t in self.columns]
        self.pfixed = [column_class(t)[1] >> 16 for _, t in self.columns]
        self.complex = [is_complex(t) for _, t in self.columns]
        self.sfixed = [column_class(t)[1] & 0xFFFF for _, t in self.static_columns]
    def col_index(self, name):
        name = name if isinstance(name, bytes) else name.encode()
        return [n for n, _ in self.columns].index(name)

def _part_dt(dt):
    if dt is None: return b"\x80"
    return struct.pack(">qI", dt[0], dt[1])

class Builder:
    def __init__(self, schema: Schema, stats=(TIMESTAMP_EPOCH, DELETION_TIME_EPOCH, 0), column_index_size=65536):
        self.s = schema; self.min_ts, self.min_ldt, self.min_ttl = stats; self.cis = column_index_size

    def clust_values(self, ck):
        out = bytearray(); off = 0
        while off < len(ck):
            limit = min(len(ck), off + 32); header = 0
            for i in range(off, limit):
                if ck[i] is None: header |= 1 << ((i % 32) * 2 + 1)
                elif len(ck[i]) == 0: header |= 1 << ((i % 32) * 2)
            out += vint(header)
            for i in range(off, limit):
                v = ck[i]
                if v is None or len(v) == 0: continue
                if not self.s.cfixed[i]: out += vint(len(v))
                out += v
            off = limit
        return bytes(out)
    def clust_prefix(self, kind, ck):
        return bytes([kind]) + (b"" if kind == K_CLUSTERING else struct.pack(">H", len(ck))) + self.clust_values(ck)
    def delta_dt(self, dt): return vint(dt[0] - self.min_ts) + i32s(dt[1] - self.min_ldt)

    def unfiltered(self, u, prev_size):
        if isinstance(u, Marker):
            head = bytes([0x02, u.kind]) + struct.pack(">H", len(u.ck)) + self.clust_values(u.ck)
            if u.kind in (K_EXCL_END_INCL_START, K_INCL_END_EXCL_START): body = self.delta_dt(u.close) + self.delta_dt(u.open)
            elif u.kind in (K_INCL_START, K_EXCL_START): body = self.delta_dt(u.open)
            else: body = self.delta_dt(u.close)
            return head + vint(len(body) + len(vint(prev_size))) + vint(prev_size) + body
        return self.row(u, prev_size, self.s.columns, self.s.vfixed, False)

    def row(self, u, prev_size, columns, vfixed, static):
        flags = 0x80 if static else 0            # static rows carry the extension byte (IS_STATIC), UnfilteredSerializer.java:109-114,755-758
        if u.ts != NO_TS: flags |= 0x04
        if u.ttl: flags |= 0x08
        if u.deletion is not None: flags |= 0x10
        cells = sorted(u.cells, key=lambda c: c.col)                       # (stable: the cells of a multi-cell column keep their path order)
        cdel = {} if static else u.complex_deletions
        present = sorted({c.col for c in cells} | set(cdel))
        if cdel: flags |= 0x40                                              # HAS_COMPLEX_DELETION
        if len(present) == len(columns): flags |= 0x20
        body = bytearray()
        if flags & 0x04: body += vint(u.ts - self.min_ts)
        if flags & 0x08: body += i32s(u.ttl - self.min_ttl) + i32s(u.ldt - self.min_ldt)
        if flags & 0x10: body += self.delta_dt(u.deletion)
        if not flags & 0x20:
            missing = (1 << len(columns)) - 1
            for col in present: missing &= ~(1 << col)
            body += vint(missing)
        cx = (lambda col: (not static) and self.s.complex[col])
        def cx_head(col):                                                   # UnfilteredSerializer.writeComplexColumn :271-280
            out = bytearray()
            if flags & 0x40: out += self.delta_dt(cdel.get(col, (NO_TS, NO_DELETION_TIME)))
            out += vint(sum(1 for c in cells if c.col == col))
            return out
        order = []                                                          # (column, cell or None) in serialisation order
        for col in present:
            mine = [c for c in cells if c.col == col]
            if cx(col): order.append((col, None)); order += [(col, c) for c in mine]
            else: order += [(col, c) for c in mine]
        for col, c in order:
            if c is None: body += cx_head(col); continue
            deleted = c.ldt != NO_DELETION_TIME and c.ttl == 0; expiring = c.ttl != 0
            use_ts = u.ts != NO_TS and c.ts == u.ts
            use_ttl = expiring and u.ttl != 0 and c.ttl == u.ttl and c.ldt == u.ldt
            cf = (0 if c.value else 0x04) | (0x01 if deleted else (0x02 if expiring else 0)) | (0x08 if use_ts else 0) | (0x10 if use_ttl else 0)
            body.append(cf)
            if not use_ts: body += vint(c.ts - self.min_ts)
            if (deleted or expiring) and not use_ttl: body += i32s(c.ldt - self.min_ldt)
            if expiring and not use_ttl: body += i32s(c.ttl - self.min_ttl)
            if cx(col):                                                     # Cell.Serializer.serialize :300-301: the path precedes the value
                if not self.s.pfixed[col]: body += vint(len(c.path))
                body += c.path
            if c.value:
                if not vfixed[c.col]: body += vint(len(c.value))
                body += c.value
        head = bytes([flags, 0x01]) if static else bytes([flags]) + self.clust_values(u.ck)
        return head + vint(len(body) + len(vint(prev_size))) + vint(prev_size) + bytes(body)

    def build(self, partitions, chunk_length=16384, compressor=O.COMP_LZ4, generation=0):
        parts = sorted(partitions, key=lambda p: (O.token(p.key), p.key))
        data = bytearray(); index = bytearray(); summary = []
        for pi, p in enumerate(parts):
            start = len(data)
            if pi % getattr(self, 'summary_interval', 3) == 0: summary.append(len(index))      # Summary.db sample (tiny interval: many anchors even in small tables)
            data += struct.pack(">H", len(p.key)) + p.key + _part_dt(p.deletion)
            if self.s.static_columns:              # SortedTablePartitionWriter.addStaticRow :117-126: always present when the header has static columns
                data += self.row(p.static if p.static is not None else Row(()), 0, self.s.static_columns, self.s.sfixed, True)
            header_len = len(data) - start
            prev_start = 0; infos = []; first = None; block_start = 0; open_marker = None; last = None
            for u in p.unfiltereds:
                pos = len(data) - start
                kind = u.kind if isinstance(u, Marker) else K_CLUSTERING
                if first is None: first = (kind, u.ck); block_start = pos
                data += self.unfiltered(u, pos - prev_start)
                prev_start = pos; last = (kind, u.ck)
                if isinstance(u, Marker): open_marker = u.open if u.kind in (K_INCL_START, K_EXCL_START, K_EXCL_END_INCL_START, K_INCL_END_EXCL_START) else None
                if len(data) - start - block_start >= self.cis:
                    infos.append(self._index_info(first, last, block_start, len(data) - start - block_start, open_marker)); first = None
            data.append(0x01)
            if p.unfiltereds and first is not None:
                infos.append(self._index_info(first, last, block_start, len(data) - start - block_start, open_marker))   # width includes the end-of-partition byte (finish() :217-243)
            index += struct.pack(">H", len(p.key)) + p.key + vint(start)
            if len(infos) > 1:
                size = len(vint(header_len)) + len(_part_dt(p.deletion)) + len(vint(len(infos))) + sum(map(len, infos)) + 4 * len(infos)
                index += vint(size) + vint(header_len) + _part_dt(p.deletion) + vint(len(infos)) + b"".join(infos)
                off = 0
                for ii in infos: index += struct.pack(">i", off); off += len(ii)
            else: index += vint(0)
        data = bytes(data)
        image = bytearray(); offs = []
        for i in range(0, len(data), chunk_length):
            c = O.chunk_compress(compressor, data[i:i + chunk_length]); offs.append(len(image)); image += c + struct.pack(">I", O.crc32(c))
        name = {O.COMP_LZ4: "LZ4Compressor", O.COMP_SNAPPY: "SnappyCompressor"}[compressor]
        meta = CompressionMetadata(name, chunk_length, 0x7FFFFFFF, len(data), offs)
        stats = (self.min_ts, self.min_ldt, self.min_ttl)
        t = SSTable(bytes(image), bytes(index), meta, stats, stats, self.s.clustering_types, self.s.columns, self.s.static_columns, generation=generation)
        t.uncompressed = data
        import numpy as _np
        t.summary_positions = _np.asarray(summary, dtype=_np.uint64)
        return t
    def _index_info(self, first, last, offset, width, open_marker):
        zz = ((width - 65536) << 1) ^ ((width - 65536) >> 63)
        return self.clust_prefix(*first) + self.clust_prefix(*last) + vint(offset) + vint(zz) + (b"\x00" if open_marker is None else b"\x01" + _part_dt(open_marker))

# ---- reading back (for semantic checks): uncompressed stream -> logical partitions ------------------------------------------
def decode_stream(schema: Schema, data: bytes, stats):
    min_ts, min_ldt, min_ttl = stats
    p = 0; parts = []
    def rv():
        nonlocal p
        f = data[p]
        if f < 0x80: p += 1; return f
        extra = 8 if f == 0xFF else (8 - (f ^ 0xFF).bit_length()); v = f & (0xFF >> extra)
        for i in range(extra): v = (v << 8) | data[p + 1 + i]
        p += 1 + extra; return v
    def ri32():
        v = rv()
        return v - (1 << 64) if v >= (1 << 63) else v
    def rclust(n):
        nonlocal p
        out = []; off = 0
        while off < n:
            limit = min(n, off + 32); header = rv()
            for i in range(off, limit):
                if (header >> ((i % 32) * 2 + 1)) & 1: out.append(None); continue
                if (header >> ((i % 32) * 2)) & 1: out.append(b""); continue
                ln = schema.cfixed[i] or rv()
                out.append(data[p:p + ln]); p += ln
            off = limit
        return tuple(out)
    def rdt(): return (rv() + min_ts, ri32() + min_ldt)
    def rrow(flags, ck, columns, vfixed):
        nonlocal p
        rv(); rv()
        r = Row(ck)
        if flags & 0x04: r.ts = rv() + min_ts
        if flags & 0x08: r.ttl = ri32() + min_ttl; r.ldt = ri32() + min_ldt
        if flags & 0x10: r.deletion = rdt()
        missing = 0 if flags & 0x20 else rv()
        for ci in range(len(columns)):
            if (missing >> ci) & 1: continue
            cf = data[p]; p += 1
            ts = r.ts if cf & 0x08 else rv() + min_ts
            ldt = r.ldt if cf & 0x10 else ((ri32() + min_ldt) if cf & 0x03 else NO_DELETION_TIME)
            ttl = r.ttl if cf & 0x10 else ((ri32() + min_ttl) if cf & 0x02 else 0)
            val = b""
            if not cf & 0x04:
                ln = vfixed[ci] or rv(); val = data[p:p + ln]; p += ln
            r.cells.append(Cell(ci, ts, val, ttl, ldt))
        return r
    while p < len(data):
        (kl,) = struct.unpack_from(">H", data, p); p += 2; key = data[p:p + kl]; p += kl
        if data[p] == 0x80: pdel = None; p += 1
        else: pdel = struct.unpack_from(">qI", data, p); p += 12
        static = None
        if schema.static_columns:
            flags, ext = data[p], data[p + 1]; p += 2
            assert flags & 0x80 and ext == 0x01 and not flags & 0x03, "static row flags"
            static = rrow(flags, (), schema.static_columns, schema.sfixed)
            if static.ts == NO_TS and static.deletion is None and not static.cells: static = None
        us = []
        while True:
            flags = data[p]; p += 1
            if flags & 1: break
            if flags & 2:
                kind = data[p]; p += 1; (n,) = struct.unpack_from(">H", data, p); p += 2; ck = rclust(n); rv(); rv()
                if kind in (K_EXCL_END_INCL_START, K_INCL_END_EXCL_START): c = rdt(); o = rdt(); us.append(Marker(kind, ck, c, o))
                elif kind in (K_INCL_START, K_EXCL_START): us.append(Marker(kind, ck, None, rdt()))
                else: us.append(Marker(kind, ck, rdt(), None))
                continue
            ck = rclust(len(schema.clustering_types))
            us.append(rrow(flags, ck, schema.columns, schema.vfixed))
        parts.append(Partition(key, us, pdel, static))
    return parts
