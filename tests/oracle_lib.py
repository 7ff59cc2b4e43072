"""ctypes access to the CPU oracle (oracle/_build/liboracle.so). TEST INFRASTRUCTURE ONLY."""
import ctypes as C, os, subprocess, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

def build():
    d = os.path.join(ROOT, "oracle")
    out = os.path.join(d, "_build", "liboracle.so")
    srcs = glob.glob(os.path.join(d, "*.cc")) + glob.glob(os.path.join(d, "*.h")) + [os.path.join(ROOT, "include", "b200c.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["make", "-C", d, "-s"])
    return out

def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p = C.c_char_p
        L.orc_vint_size.argtypes = [C.c_uint64]; L.orc_vint_size.restype = C.c_int
        L.orc_vint_write.argtypes = [C.c_void_p, C.c_uint64]; L.orc_vint_write.restype = C.c_int
        L.orc_vint_read.argtypes = [u8p, C.c_int, C.POINTER(C.c_uint64)]; L.orc_vint_read.restype = C.c_int
        L.orc_crc32.argtypes = [C.c_uint32, u8p, C.c_uint64]; L.orc_crc32.restype = C.c_uint32
        L.orc_crc32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]; L.orc_crc32_combine.restype = C.c_uint32
        L.orc_murmur3_token.argtypes = [u8p, C.c_uint64]; L.orc_murmur3_token.restype = C.c_int64
        L.orc_murmur3_x64_128.argtypes = [u8p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
        for f in ("orc_lz4_compress_block", "orc_lz4_decompress_block"):
            getattr(L, f).argtypes = [u8p, C.c_int, C.c_void_p, C.c_int]; getattr(L, f).restype = C.c_int
        L.orc_lz4_compress_bound.argtypes = [C.c_int]; L.orc_lz4_compress_bound.restype = C.c_int
        L.orc_snappy_max_compressed_length.argtypes = [C.c_int]; L.orc_snappy_max_compressed_length.restype = C.c_int
        L.orc_snappy_compress.argtypes = [u8p, C.c_int, C.c_void_p]; L.orc_snappy_compress.restype = C.c_int
        L.orc_snappy_decompress.argtypes = [u8p, C.c_int, C.c_void_p, C.c_int]; L.orc_snappy_decompress.restype = C.c_int
        L.orc_chunk_max_compressed.argtypes = [C.c_int, C.c_int]; L.orc_chunk_max_compressed.restype = C.c_int
        L.orc_chunk_compress.argtypes = [C.c_int, u8p, C.c_int, C.c_void_p]; L.orc_chunk_compress.restype = C.c_int
        L.orc_chunk_decompress.argtypes = [C.c_int, u8p, C.c_int, C.c_void_p, C.c_int]; L.orc_chunk_decompress.restype = C.c_int
        _LIB = L
    return _LIB

COMP_LZ4, COMP_SNAPPY, COMP_SNAPPY15 = 1, 2, 3

def lz4_compress(b: bytes) -> bytes:
    L = lib(); cap = L.orc_lz4_compress_bound(len(b)); out = C.create_string_buffer(cap)
    n = L.orc_lz4_compress_block(b, len(b), out, cap); assert n > 0
    return out.raw[:n]

def lz4_decompress(b: bytes, ulen: int) -> bytes:
    L = lib(); out = C.create_string_buffer(max(ulen, 1))
    n = L.orc_lz4_decompress_block(b, len(b), out, ulen)
    if n < 0: raise ValueError("malformed lz4 block")
    return out.raw[:n]

def chunk_compress(comp: int, b: bytes) -> bytes:
    L = lib(); out = C.create_string_buffer(L.orc_chunk_max_compressed(comp, len(b)) + 8)
    n = L.orc_chunk_compress(comp, b, len(b), out); assert n > 0
    return out.raw[:n]

def chunk_decompress(comp: int, b: bytes, cap: int) -> bytes:
    L = lib(); out = C.create_string_buffer(max(cap, 1))
    n = L.orc_chunk_decompress(comp, b, len(b), out, cap)
    if n < 0: raise ValueError("malformed chunk")
    return out.raw[:n]

def crc32(b: bytes, crc: int = 0) -> int:
    return lib().orc_crc32(crc, b, len(b))

def vint(v: int) -> bytes:
    out = C.create_string_buffer(9); n = lib().orc_vint_write(out, v); return out.raw[:n]

def token(key: bytes) -> int:
    return lib().orc_murmur3_token(key, len(key))

class OracleEngine:
    """orc_compact: the CPU oracle behind the same manifest/result structs as b200c_compact (TEST INFRASTRUCTURE)."""
    needs_lib_bound = False
    def __call__(self, manifest, result):
        from cassandra_b200 import native
        L = lib()
        L.orc_compact.restype = C.c_int
        err = C.create_string_buffer(256)
        rc = L.orc_compact(C.byref(manifest), C.byref(result), err, 256)
        if rc == native.ECORRUPT: raise native.CorruptSSTableError(rc, err.value.decode(), result.corruption)
        if rc == native.EUNSUPPORTED: raise native.UnsupportedError(rc, err.value.decode())
        if rc != 0: raise native.B200CError(rc, err.value.decode())
