"""Mirror of org.apache.cassandra.io.compress for the B200 engine.

ICompressor (S/io/compress/ICompressor.java:28-86): initialCompressedBufferLength / compress / uncompress, stateless singletons
created by `create(options)` (S/schema/CompressionParams.java:266-283). The persisted class simple name stays
`LZ4Compressor` / `SnappyCompressor` so stock nodes can read the files (S/io/compress/CompressionMetadata.java:379).
CompressionMetadata: CompressionInfo.db reader/writer (S/io/compress/CompressionMetadata.java:113-160,375-431).
"""
import struct
from .. import native

class ICompressor:
    compressor_id = native.COMP_NONE
    simple_name = "NoopCompressor"
    def __init__(self, ctx): self.ctx = ctx
    def initial_compressed_buffer_length(self, chunk_length):            # ICompressor.java:30
        return native.lib().b200c_initial_compressed_buffer_length(self.compressor_id, chunk_length)
    def compress(self, data: bytes) -> bytes:                            # ICompressor.java:51  (ByteBuffer in -> out)
        import ctypes as C
        cap = self.initial_compressed_buffer_length(max(len(data), 1)) + 8
        out = C.create_string_buffer(cap)
        n = native.lib().b200c_compress(self.ctx.handle, self.compressor_id, data, len(data), out, cap)
        self.ctx.check(n)
        return out.raw[:n]
    def uncompress(self, data: bytes, max_len=65536) -> bytes:           # ICompressor.java:37,58; IOException on malformed input
        import ctypes as C
        out = C.create_string_buffer(max(max_len, 1))
        n = native.lib().b200c_uncompress(self.ctx.handle, self.compressor_id, data, len(data), out, max_len)
        self.ctx.check(n)
        return out.raw[:n]
    def supported_options(self): return set()

class LZ4Compressor(ICompressor):
    """S/io/compress/LZ4Compressor.java — fast mode only (lz4_compressor_type = fast, :48,102-106)."""
    compressor_id = native.COMP_LZ4
    simple_name = "LZ4Compressor"
    @classmethod
    def create(cls, ctx, options=None):
        options = options or {}
        if options.get("lz4_compressor_type", "fast") != "fast":
            raise native.UnsupportedError(native.EUNSUPPORTED, "only lz4_compressor_type=fast is implemented on the GPU")
        return cls(ctx)
    def supported_options(self): return {"lz4_high_compressor_level", "lz4_compressor_type"}

class SnappyCompressor(ICompressor):
    """S/io/compress/SnappyCompressor.java — raw snappy. Byte parity with snappy-java is UNPINNED (no golden)."""
    compressor_id = native.COMP_SNAPPY
    simple_name = "SnappyCompressor"
    @classmethod
    def create(cls, ctx, options=None): return cls(ctx)

COMPRESSOR_IDS = {"LZ4Compressor": native.COMP_LZ4, "SnappyCompressor": native.COMP_SNAPPY}
COMPRESSOR_NAMES = {v: k for k, v in COMPRESSOR_IDS.items()}

class CompressionMetadata:
    """CompressionInfo.db. Layout: UTF(simple class name) | i32 nOpts | (UTF k, UTF v)* | i32 chunkLength |
    i32 maxCompressedLength (version >= na) | i64 dataLength | i32 nChunks | i64 offset x n — all big-endian."""
    def __init__(self, compressor_name, chunk_length, max_compressed_length, data_length, chunk_offsets, options=None):
        self.compressor_name = compressor_name; self.chunk_length = chunk_length
        self.max_compressed_length = max_compressed_length; self.data_length = data_length
        self.chunk_offsets = list(chunk_offsets); self.options = dict(options or {})
    @property
    def compressor_id(self): return COMPRESSOR_IDS[self.compressor_name]
    @classmethod
    def parse(cls, b: bytes, has_max_compressed_length=True):
        p = 0
        (n,) = struct.unpack_from(">H", b, p); p += 2; name = b[p:p + n].decode(); p += n
        (nopt,) = struct.unpack_from(">i", b, p); p += 4; opts = {}
        for _ in range(nopt):
            (kl,) = struct.unpack_from(">H", b, p); p += 2; k = b[p:p + kl].decode(); p += kl
            (vl,) = struct.unpack_from(">H", b, p); p += 2; v = b[p:p + vl].decode(); p += vl
            opts[k] = v
        (cl,) = struct.unpack_from(">i", b, p); p += 4
        mcl = native.INT32_MAX
        if has_max_compressed_length: (mcl,) = struct.unpack_from(">i", b, p); p += 4
        (dl, nc) = struct.unpack_from(">qi", b, p); p += 12
        offs = list(struct.unpack_from(">%dq" % nc, b, p))
        return cls(name, cl, mcl, dl, offs, opts)
    def serialize(self) -> bytes:                                         # CompressionMetadata.Writer.writeHeader :375-398 + doPrepare :423-431
        name = self.compressor_name.encode()
        out = [struct.pack(">H", len(name)), name, struct.pack(">i", len(self.options))]
        for k, v in self.options.items():
            kb, vb = k.encode(), v.encode(); out += [struct.pack(">H", len(kb)), kb, struct.pack(">H", len(vb)), vb]
        out.append(struct.pack(">iiqi", self.chunk_length, self.max_compressed_length, self.data_length, len(self.chunk_offsets)))
        out.append(struct.pack(">%dq" % len(self.chunk_offsets), *self.chunk_offsets))
        return b"".join(out)

def write_compressed(ctx, compressor: ICompressor, stream: bytes, chunk_length=16384, max_compressed_length=native.INT32_MAX):
    """CompressedSequentialWriter over a whole uncompressed stream (S/io/compress/CompressedSequentialWriter.java:140-206):
    -> (Data.db bytes, CompressionInfo.db bytes, Digest.crc32 text)."""
    image, offs, digest = ctx.compress_chunks(compressor.compressor_id, stream, chunk_length, max_compressed_length)
    meta = CompressionMetadata(compressor.simple_name, chunk_length, max_compressed_length, len(stream), offs)
    return image, meta.serialize(), str(digest)

def read_compressed(ctx, data_db: bytes, compression_info: bytes, verify_crc=True) -> bytes:
    """CompressedChunkReader over a whole file (S/io/util/CompressedChunkReader.java:103-173) -> uncompressed stream."""
    meta = CompressionMetadata.parse(compression_info)
    return ctx.decompress_chunks(meta.compressor_id, data_db, meta.chunk_offsets, meta.data_length, meta.chunk_length,
                                 meta.max_compressed_length, verify_crc)
