"""Minimal pure-Python readers for SSTable component files used by tests (CompressionInfo.db, Data.db chunking).
Follows S/io/compress/CompressionMetadata.java:375-398,423-431 (layout) — test helper, not product code."""
import struct, os, glob

def read_compression_info(path):
    b = open(path, "rb").read(); p = 0
    (n,) = struct.unpack_from(">H", b, p); p += 2
    name = b[p:p + n].decode(); p += n
    (nopt,) = struct.unpack_from(">i", b, p); p += 4
    opts = {}
    for _ in range(nopt):
        (kl,) = struct.unpack_from(">H", b, p); p += 2; k = b[p:p + kl].decode(); p += kl
        (vl,) = struct.unpack_from(">H", b, p); p += 2; v = b[p:p + vl].decode(); p += vl
        opts[k] = v
    (chunk_len,) = struct.unpack_from(">i", b, p); p += 4
    max_clen = None
    # 'maxCompressedLength' exists from version na on (CompressionMetadata.java:113-116); detect by remaining size
    rest = len(b) - p
    # try with max_compressed_length
    (mc, dl, nc) = struct.unpack_from(">iqi", b, p)
    if rest == 4 + 8 + 4 + 8 * nc:
        max_clen = mc; p += 16
    else:
        (dl, nc) = struct.unpack_from(">qi", b, p); p += 12
        assert rest == 8 + 4 + 8 * nc, (rest, nc)
    offsets = list(struct.unpack_from(">%dq" % nc, b, p))
    return dict(compressor=name, options=opts, chunk_length=chunk_len, max_compressed_length=max_clen,
                data_length=dl, offsets=offsets)

def split_chunks(data: bytes, info):
    """-> list of (compressed_bytes, crc_be_u32, uncompressed_len)"""
    offs = info["offsets"]; out = []
    for i, o in enumerate(offs):
        end = (offs[i + 1] if i + 1 < len(offs) else len(data)) - 4
        (crc,) = struct.unpack_from(">I", data, end)
        ulen = min(info["chunk_length"], info["data_length"] - i * info["chunk_length"])
        out.append((data[o:end], crc, ulen))
    return out

def find_tables(golden_dir, versions=None):
    res = []
    for f in sorted(glob.glob(os.path.join(golden_dir, "*", "legacy_tables", "*", "*-big-Data.db"))):
        v = f.split(os.sep)[-4]
        if versions and v not in versions: continue
        base = f[:-len("Data.db")]
        if os.path.exists(base + "CompressionInfo.db"):
            res.append(base)
    return res
