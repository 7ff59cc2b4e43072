"""K1's two-pass LZ4 decoder (cassandra_b200/csrc/lz4_batch.cuh) on the CPU: the walk (plain C++) and the warp-per-chunk copy (on the warp emulator
of tests/native/warp_emu.h) must reproduce the oracle's decoder on valid blocks and agree with it on which damaged blocks are refused."""
import ctypes as C, os, random, shutil, subprocess, pytest
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

@pytest.fixture(scope="module")
def lib():
    if shutil.which("g++") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"): pytest.skip("needs g++ and the CUDA headers")
    out = os.path.join(ROOT, "tests", "native", "_build", "liblz4batch.so")
    srcs = [os.path.join(ROOT, "tests", "native", "lz4_batch_host.cc")]
    deps = srcs + [os.path.join(ROOT, "tests", "native", "warp_emu.h")] + [os.path.join(ROOT, "cassandra_b200", "csrc", f) for f in ("lz4_batch.cuh", "lz4_thread.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        r = subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-fno-strict-aliasing", "-I/usr/local/cuda/include", "-Wno-attributes", "-Wno-unknown-pragmas",
                            "-o", out] + srcs, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
    L = C.CDLL(out); L.batch_decompress.restype = C.c_int; L.batch_decompress.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    return L

def decode(L, block, ulen):
    dst = C.create_string_buffer(ulen + 16); ns = C.c_int(0)
    got = L.batch_decompress(block + bytes(16), len(block), dst, ulen, C.byref(ns))
    return got, dst.raw[:max(ulen, 0)], ns.value

def shapes(rng):
    alphabet = [bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 9))) for _ in range(30)]
    for it in range(120):
        n = rng.choice([0, 1, 5, 12, 13, 14, 31, 32, 33, 64, 100, 255, 1000, 4096, 16384, 16384, 16384, 40000, 65536])
        kind = rng.random()
        if kind < 0.15: d = bytes(rng.getrandbits(8) for _ in range(n))                      # one long literal run (copied by the whole warp)
        elif kind < 0.5: d = b"".join(rng.choice(alphabet) for _ in range(n))[:n]             # many short sequences, near and far matches
        elif kind < 0.6: d = bytes(rng.choice(b"ab\x00") for _ in range(n))
        elif kind < 0.75: d = (bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 40))) * (n + 1))[:n]     # overlapping matches (offset < length), very long
        elif kind < 0.85: d = bytes(n)                                                          # offset 1
        else:                                                                                   # matches that depend on matches of the same step
            seed = bytes(rng.getrandbits(8) for _ in range(rng.randint(4, 12))); out = bytearray(seed)
            while len(out) < n: k = rng.randint(4, 12); s = rng.randint(0, len(out) - 1); out += out[s:s + k] + bytes([rng.getrandbits(8)])
            d = bytes(out[:n])
        yield it, d

def test_valid_blocks_decode_exactly(lib):
    for it, d in shapes(random.Random(0xBA7C4)):
        block = O.lz4_compress(d)
        got, out, ns = decode(lib, block, len(d))
        assert got == len(d) and out == d, (it, len(d), got, ns)

def test_damaged_blocks_are_refused_like_the_oracle(lib):
    rng = random.Random(0xD4A6ED)
    refused = 0
    for it, d in shapes(rng):
        if not d: continue
        block = bytearray(O.lz4_compress(d))
        for m in range(6):
            bad = bytearray(block)
            if m & 1: bad = bad[:rng.randrange(len(bad))]
            else:
                for _ in range(rng.randint(1, 3)): bad[rng.randrange(len(bad))] ^= 1 << rng.randrange(8)
            try: want = O.lz4_decompress(bytes(bad), len(d))
            except ValueError: want = None
            if want is not None and len(want) != len(d): want = None                           # the chunk reader also refuses a short decode
            got, out, ns = decode(lib, bytes(bad), len(d))
            if want is None: assert got == -1, (it, m, got); refused += 1
            else: assert got == len(d) and out == want, (it, m)
        got, _, _ = decode(lib, bytes(block), len(d) - 1)                                      # wrong expected size
        assert got == -1
    assert refused > 100
