"""Device-resident micro-benchmark of the chunk codec kernels (development aid; bench.py is the contract bench)."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cassandra_b200 import native

def sstable_like(nbytes, seed=1):
    rng = np.random.default_rng(seed)
    nrows = nbytes // 27
    rows = np.zeros((nrows, 27), dtype=np.uint8)
    rows[:, 0] = 0x24
    ck = np.cumsum(rng.integers(1, 1 << 40, size=nrows, dtype=np.int64)).astype(">i8")
    rows[:, 1:9] = ck.view(np.uint8).reshape(-1, 8)
    rows[:, 9] = 16; rows[:, 10] = 27
    ts = rng.integers(0, 1 << 30, size=nrows, dtype=np.int64).astype(">i4")
    rows[:, 11] = 0xE0 | (rng.integers(0, 16, size=nrows)); rows[:, 12:16] = ts.view(np.uint8).reshape(-1, 4)[:, :4]
    rows[:, 16] = 0x08
    val = rng.integers(0, 1 << 20, size=nrows, dtype=np.int64).astype(">i8")
    rows[:, 17:25] = val.view(np.uint8).reshape(-1, 8)
    rows[:, 25] = 0x01; rows[:, 26] = 0
    return rows.reshape(-1).tobytes()

def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    comp = native.COMP_LZ4
    stream = sstable_like(mb << 20)
    n = len(stream); L = native.lib()
    with native.Context(0) as ctx:
        h = ctx.handle
        d_in = C.c_void_p(); d_out = C.c_void_p(); d_offs = C.c_void_p(); d_back = C.c_void_p()
        cap = L.b200c_compress_bound(comp, n, 16384); nch = L.b200c_chunk_count(n, 16384)
        for p, sz in ((d_in, n), (d_out, cap), (d_offs, 8 * (nch + 1)), (d_back, n)):
            ctx.check(L.b200c_dev_alloc(h, sz, C.byref(p)))
        ctx.check(L.b200c_memcpy_h2d(h, d_in, stream, n))
        out_len = C.c_uint64(); dig = C.c_uint32()
        for it in range(4):
            ctx.check(L.b200c_compress_chunks(h, comp, d_in, n, 16384, native.INT32_MAX, d_out, cap, C.byref(out_len), d_offs, C.byref(dig), 1))
            ms = ctx.last_kernel_ms
            print("compress   %d MiB -> %.1f MiB  %.3f ms  %.2f GB/s (uncompressed in)" % (mb, out_len.value / 2**20, ms, n / ms / 1e6))
        for it in range(4):
            ctx.check(L.b200c_decompress_chunks(h, comp, d_out, out_len.value, d_offs, nch, 16384, native.INT32_MAX, n, d_back, 1, None, 1))
            ms = ctx.last_kernel_ms
            print("decompress %.1f MiB -> %d MiB  %.3f ms  %.2f GB/s (uncompressed out)" % (out_len.value / 2**20, mb, ms, n / ms / 1e6))
        back = np.empty(n, dtype=np.uint8)
        ctx.check(L.b200c_memcpy_d2h(h, back.ctypes.data, d_back, n))
        assert back.tobytes() == stream
        print("roundtrip ok, ratio %.3f" % (out_len.value / n))
if __name__ == "__main__":
    main()
