"""Seeded generator of synthetic big-format `oa` input SSTables. Self-contained C++ (synth/synth.cc);
never imports oracle/. The caller supplies the chunk compressor (tests: CPU oracle; bench: the engine's own K5)."""
import ctypes as C, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NOW_IN_SEC = 1700000000
GC_GRACE = 864000
MARSHAL = "org.apache.cassandra.db.marshal."
SCHEMAS = {
    0: dict(clustering=[MARSHAL + "LongType"], columns=[(b"commentid", MARSHAL + "LongType")]),
    1: dict(clustering=[MARSHAL + "TimestampType"], columns=[(b"tag", MARSHAL + "UTF8Type"), (b"v1", MARSHAL + "DoubleType"), (b"v2", MARSHAL + "DoubleType")]),
}

class Config(C.Structure):
    _fields_ = [("schema", C.c_int32), ("sstable", C.c_int32), ("nsstables", C.c_int32), ("rows_per_partition", C.c_int32),
                ("seed", C.c_uint64), ("universe", C.c_uint64), ("p", C.c_double), ("column_index_size", C.c_int32),
                ("threads", C.c_int32), ("band_count", C.c_int32), ("l0_count", C.c_int32)]
class Result(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_len", C.c_uint64), ("index", C.c_void_p), ("index_len", C.c_uint64),
                ("partitions", C.c_uint64), ("rows", C.c_uint64), ("min_timestamp", C.c_int64),
                ("min_local_deletion_time", C.c_int64), ("min_ttl", C.c_int32), ("_pad", C.c_int32),
                ("summary", C.c_void_p), ("nsummary", C.c_uint64)]

def build():
    out = os.path.join(_HERE, "_build", "libsynth.so")
    src = os.path.join(_HERE, "synth.cc")
    if not os.path.exists(out) or os.path.getmtime(src) > os.path.getmtime(out):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return out

def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.synth_generate.argtypes = [C.POINTER(Config), C.POINTER(Result)]; L.synth_generate.restype = C.c_int
        L.synth_free.argtypes = [C.POINTER(Result)]
        L.synth_universe_for.argtypes = [C.c_int, C.c_uint64, C.c_double, C.c_int]; L.synth_universe_for.restype = C.c_uint64
        _LIB = L
    return _LIB

def universe_for(schema, target_bytes, p, rows_per_partition=1000):
    return lib().synth_universe_for(schema, target_bytes, p, rows_per_partition)

def generate_raw(schema, sstable, nsstables, seed, universe, p=0.5, rows_per_partition=1000, column_index_size=65536, threads=0,
                 band_count=0, l0_count=0):
    """-> dict(stream: np.uint8 array (uncompressed Data stream), index: bytes, partitions, rows, stats=(min_ts, min_ldt, min_ttl))"""
    cfg = Config(schema, sstable, nsstables, rows_per_partition, seed, universe, p, column_index_size, threads, band_count, l0_count)
    res = Result()
    if lib().synth_generate(C.byref(cfg), C.byref(res)) != 0: raise MemoryError("synth_generate failed")
    try:
        stream = np.ctypeslib.as_array(C.cast(res.data, C.POINTER(C.c_uint8)), shape=(res.data_len,)).copy() if res.data_len else np.zeros(0, np.uint8)
        index = C.string_at(res.index, res.index_len)
        summary = np.ctypeslib.as_array(C.cast(res.summary, C.POINTER(C.c_uint64)), shape=(res.nsummary,)).copy() if res.nsummary else np.zeros(0, np.uint64)
        return dict(stream=stream, index=index, partitions=res.partitions, rows=res.rows, summary=summary,
                    stats=(res.min_timestamp, res.min_local_deletion_time, res.min_ttl))
    finally:
        lib().synth_free(C.byref(res))

def make_sstable(raw, schema, compress, compressor_name="LZ4Compressor", chunk_length=16384, generation=0, level=0):
    """compress(stream ndarray/bytes, chunk_length) -> (Data.db image bytes-like, [chunk offsets]). Returns an io.sstable.SSTable."""
    from cassandra_b200.io.sstable import SSTable
    from cassandra_b200.io.compress import CompressionMetadata
    image, offs = compress(raw["stream"], chunk_length)
    meta = CompressionMetadata(compressor_name, chunk_length, 0x7FFFFFFF, len(raw["stream"]), offs)
    sc = SCHEMAS[schema]
    t = SSTable(image, raw["index"], meta, raw["stats"], raw["stats"], sc["clustering"], sc["columns"], generation=generation, level=level)
    t.partitions = raw["partitions"]; t.rows = raw["rows"]; t.summary_positions = raw.get("summary")
    return t
