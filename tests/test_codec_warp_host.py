"""The warp-per-chunk compressors of K5 (cassandra_b200/csrc/lz4.cuh, snappy.cuh) on the CPU: tests/native/codec_warp_host.cc compiles the very
device source with g++ and runs it on a 32-fiber warp emulator (tests/native/warp_emu.h: every warp intrinsic is a lockstep rendezvous, a
collective that not all lanes reach aborts). Output must equal the oracle's / liblz4's / Google snappy's byte for byte — the GPU tests prove the
same for the CUDA build; this one needs no GPU and lets the speculation logic be developed here."""
import ctypes as C, os, random, shutil, subprocess, pytest
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

@pytest.fixture(scope="module")
def warp():
    if shutil.which("g++") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"): pytest.skip("needs g++ and the CUDA headers")
    out = os.path.join(ROOT, "tests", "native", "_build", "libcodecwarp.so")
    srcs = [os.path.join(ROOT, "tests", "native", "codec_warp_host.cc")]
    deps = srcs + [os.path.join(ROOT, "tests", "native", "warp_emu.h")] + [os.path.join(ROOT, "cassandra_b200", "csrc", f) for f in ("lz4.cuh", "lz4_chain.cuh", "snappy.cuh", "snappy_chain.cuh", "common.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        r = subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-fno-strict-aliasing", "-I/usr/local/cuda/include", "-Wno-attributes", "-Wno-unknown-pragmas",
                            "-o", out] + srcs, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
    L = C.CDLL(out); L.warp_compress.restype = C.c_int; L.warp_compress.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_void_p]
    return L

def run(L, mode, data):
    out = C.create_string_buffer(len(data) + len(data) // 6 + 64)
    n = L.warp_compress(mode, data, len(data), out); assert n > 0
    return out.raw[:n]

def corpus():
    from test_oracle_codec import _corpus
    rng = random.Random(0xC0DEC)
    words = [bytes(rng.getrandbits(8) for _ in range(rng.randint(2, 12))) for _ in range(100)]
    extra = [b"".join(rng.choice(words) for _ in range(3000))[:16384], bytes(16384), bytes(rng.getrandbits(8) for _ in range(5000)), b"ab" * 4000, b"x" * 13, b"y" * 12, b"0123456789abcdef" * 700]
    return list(_corpus(random.Random(7)))[:12] + extra

@pytest.mark.parametrize("mode", [0, 1, 4, 7])
def test_lz4_warp_source_equals_the_oracle(warp, mode):
    for k, d in enumerate(corpus()):
        d = d[:16384]
        if not d: continue
        assert run(warp, mode, d) == O.lz4_compress(d), (k, len(d))

def test_snappy_warp_source_equals_googles_library(warp):
    from test_snappy_golden import vectors
    for name, data, want in vectors():
        if not data: continue
        assert run(warp, 3, data) == want, name
        if len(data) <= 32768: assert run(warp, 9, data) == want, name                           # two passes (snappy_chain.cuh)
        assert run(warp, 6, data) == want, name                                                  # chunk read in place (the L1 variant K5 launches)
        assert run(warp, 2, data) == O.chunk_compress(O.COMP_SNAPPY, data), name

def test_snappy_warp_source_random_differential(warp):
    """more shapes than the golden vectors: sizes around the 15-byte margin, the 16-attempt prologue, table-size steps and the 64 KiB fragment"""
    rng = random.Random(0x5A9)
    alphabet = [bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 9))) for _ in range(30)]
    for it in range(60):
        n = rng.choice([1, 2, 14, 15, 16, 17, 30, 31, 32, 33, 100, 255, 256, 257, 1000, 4096, 4097, 16384, 20000, 65536])
        kind = rng.random()
        if kind < 0.25: d = bytes(rng.getrandbits(8) for _ in range(n))
        elif kind < 0.6: d = b"".join(rng.choice(alphabet) for _ in range(n))[:n]
        elif kind < 0.8: d = bytes(rng.choice(b"ab\x00") for _ in range(n))
        else: d = (bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 40))) * n)[:n]
        assert run(warp, 2, d) == O.chunk_compress(O.COMP_SNAPPY, d), (it, n, "2^14")
        assert run(warp, 3, d) == O.chunk_compress(O.COMP_SNAPPY15, d), (it, n, "2^15")
        assert run(warp, 5, d) == O.chunk_compress(O.COMP_SNAPPY, d), (it, n, "2^14 in place")
        if n <= 32768:
            assert run(warp, 8, d) == O.chunk_compress(O.COMP_SNAPPY, d), (it, n, "2^14 two passes")
            assert run(warp, 9, d) == O.chunk_compress(O.COMP_SNAPPY15, d), (it, n, "2^15 two passes")

def test_lz4_chain_random_differential(warp):
    """lz4_chain.cuh (mode 7) against the oracle on the shapes that stress what it changed: incompressible data (accelerated, non-contiguous search
    windows), long runs and short periods (many same-hash positions per step, deep chains of never-inserted positions), sizes around the 13-byte
    minimum and up to the 32 KiB limit of the 15-bit links."""
    rng = random.Random(0x17C4A1)
    alphabet = [bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 9))) for _ in range(30)]
    sizes = [1, 4, 11, 12, 13, 14, 15, 16, 17, 31, 32, 33, 44, 45, 63, 64, 65, 66, 67, 100, 255, 256, 257, 1000, 4095, 4096, 4097, 8192, 16383, 16384, 20000, 32767, 32768]
    for it in range(160):
        n = rng.choice(sizes)
        kind = rng.random()
        if kind < 0.2: d = bytes(rng.getrandbits(8) for _ in range(n))
        elif kind < 0.5: d = b"".join(rng.choice(alphabet) for _ in range(n))[:n]
        elif kind < 0.65: d = bytes(rng.choice(b"ab\x00") for _ in range(n))
        elif kind < 0.8: d = (bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 40))) * n)[:n]
        elif kind < 0.9:                                                   # incompressible stretches between repeated blocks
            blk = bytes(rng.getrandbits(8) for _ in range(rng.randint(8, 300)))
            d = b"".join((bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 900))) + blk) for _ in range(n // 200 + 1))[:n]
        else: d = bytes(n)
        assert run(warp, 7, d) == O.lz4_compress(d), (it, n, kind)
