"""Deterministic inputs of the Snappy golden vectors (tests/golden/snappy/vectors.json, made by tests/golden/tools/make_snappy_vectors.py with
Google's snappy library). Shapes: empty / tiny / incompressible / long runs / short-period repeats / text / SSTable-like rows, at sizes around
the 16 KiB chunk length and the 64 KiB block and hash-table limits of the format."""
import random, struct

def inputs():
    out = []
    rng = random.Random(0x5A99)
    words = [bytes(rng.getrandbits(8) for _ in range(rng.randint(2, 12))) for _ in range(200)]
    def text(n):
        b = bytearray()
        while len(b) < n: b += rng.choice(words) + b" "
        return bytes(b[:n])
    def rows(n):                      # narrow rows like schema N: flags, 8-byte clustering, small vints, 8-byte value
        b = bytearray(); ck = 0
        while len(b) < n:
            ck += rng.randint(1, 9)
            b += bytes([0x24]) + struct.pack(">q", ck) + bytes([rng.randint(12, 20), rng.randint(20, 30)]) + struct.pack(">I", rng.getrandbits(20))[1:] + b"\x08" + struct.pack(">q", rng.getrandbits(30))
        return bytes(b[:n])
    out.append(("empty", b""))
    out.append(("one", b"x"))
    for n in (3, 15, 16, 17, 59, 60, 61, 64, 255, 256, 257):
        out.append(("text-%d" % n, text(n)))
    out.append(("random-4k", bytes(rng.getrandbits(8) for _ in range(4096))))
    out.append(("random-16k", bytes(rng.getrandbits(8) for _ in range(16384))))
    out.append(("zeros-16k", bytes(16384)))
    out.append(("zeros-64k", bytes(65536)))
    out.append(("period3-16k", (b"abc" * 6000)[:16384]))
    out.append(("period70-16k", (bytes(range(70)) * 300)[:16384]))
    out.append(("text-16k", text(16384)))
    out.append(("text-16383", text(16383)))
    out.append(("text-32k", text(32768)))
    out.append(("text-64k", text(65536)))
    out.append(("rows-16k", rows(16384)))
    out.append(("rows-4097", rows(4097)))
    out.append(("rows-64k", rows(65536)))
    out.append(("mixed-16k", (text(5000) + bytes(rng.getrandbits(8) for _ in range(3000)) + bytes(2000) + rows(6384))[:16384]))
    return out
