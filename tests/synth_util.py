"""Helpers shared by tests: synthetic inputs compressed with the CPU oracle (tests may use the oracle; bench.py may not)."""
import struct
import numpy as np
import oracle_lib as O
import synth

def oracle_compress(stream, chunk_length, comp=O.COMP_LZ4):
    """CPU oracle's CompressedSequentialWriter over a whole stream (one C call: releases the GIL)."""
    import ctypes as C
    a = np.ascontiguousarray(np.frombuffer(stream, dtype=np.uint8) if isinstance(stream, (bytes, bytearray)) else stream)
    n = len(a); nch = (n + chunk_length - 1) // chunk_length
    L = O.lib()
    L.orc_compress_stream.restype = C.c_uint64
    L.orc_compress_stream.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    out = np.empty(nch * (L.orc_chunk_max_compressed(comp, chunk_length) + 4) + 64, dtype=np.uint8); offs = np.zeros(max(nch, 1), dtype=np.uint64)
    m = L.orc_compress_stream(comp, a.ctypes.data, n, chunk_length, out.ctypes.data, offs.ctypes.data)
    return out[:m].tobytes(), [int(x) for x in offs[:nch]]

def synth_tables(schema, n, seed, universe, p=0.5, rows_per_partition=1000, comp=O.COMP_LZ4, **kw):
    name = {O.COMP_LZ4: "LZ4Compressor", O.COMP_SNAPPY: "SnappyCompressor"}[comp]
    out = []
    for s in range(n):
        raw = synth.generate_raw(schema, s, n, seed, universe, p, rows_per_partition, threads=4, **kw)
        t = synth.make_sstable(raw, schema, lambda st, cl: oracle_compress(st, cl, comp), name, generation=s)
        t.uncompressed = raw["stream"].tobytes()
        out.append(t)
    return out

def decompress_output(o, comp=O.COMP_LZ4):
    offs = o.compression.chunk_offsets
    return b"".join(O.chunk_decompress(comp, o.data[a:(b if b else len(o.data)) - 4], o.compression.chunk_length)
                    for a, b in zip(offs, offs[1:] + [None]))
