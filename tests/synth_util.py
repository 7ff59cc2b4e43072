"""Helpers shared by tests: synthetic inputs compressed with the CPU oracle (tests may use the oracle; bench.py may not)."""
import struct
import numpy as np
import oracle_lib as O
import synth

def oracle_compress(stream, chunk_length, comp=O.COMP_LZ4):
    b = stream.tobytes() if hasattr(stream, "tobytes") else bytes(stream)
    image = bytearray(); offs = []
    for i in range(0, len(b), chunk_length):
        c = O.chunk_compress(comp, b[i:i + chunk_length]); offs.append(len(image)); image += c + struct.pack(">I", O.crc32(c))
    return bytes(image), offs

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
