"""Randomised tables with multi-cell (complex) columns — a map, a set and a list beside a simple column — for the K4 host-build and GPU parity tests
(the oracle they are compared with is pinned by tests/test_oracle_complex_kats.py)."""
import random, struct
from sstable_builder import *

NOW = 1700000000
I32 = lambda v: struct.pack(">i", v)
I64 = lambda v: struct.pack(">q", v)
T0 = 1_600_000_000_000_000
SCX = Schema(["Int32Type"], [("a", "UTF8Type"), ("z", "Int32Type"), ("l", "ListType(UTF8Type)"), ("m", "MapType(UTF8Type,Int32Type)"), ("s", "SetType(LongType)")])
A, Z, L_, M, S_ = 0, 1, 2, 3, 4

def timeuuid(ts100ns, node): return struct.pack(">IHHQ", ts100ns & 0xFFFFFFFF, (ts100ns >> 32) & 0xFFFF, 0x1000 | ((ts100ns >> 48) & 0x0FFF), node)
def timeuuid_key(b):
    msb, lsb = struct.unpack(">QQ", b)
    re = ((msb << 48) | ((msb << 16) & 0xFFFF00000000) | (msb >> 32)) & ((1 << 64) - 1)
    sgn = lambda x: x - (1 << 64) if x >= 1 << 63 else x
    return (sgn(re), sgn(lsb ^ 0x0080808080808080))

def complex_tables(seed, ntables=4, nkeys=60, cis=65536, big=False):
    rng = random.Random(seed)
    keys = sorted({b"key%05d" % rng.randint(0, 10 ** 5) for _ in range(nkeys)})
    uuids = sorted({timeuuid(rng.randint(1, 1 << 40), rng.getrandbits(64)) for _ in range(12)}, key=timeuuid_key)
    tables = []
    for t in range(ntables):
        parts = []
        for key in keys:
            if rng.random() < 0.4: continue
            rows = []
            for ck in sorted({rng.randint(0, 6 if not big else 300) for _ in range(rng.randint(1, 4) if not big else 120)}):
                cells = []
                ts = lambda: T0 + t * 100 + rng.randint(0, 90)
                if rng.random() < 0.5: cells.append(Cell(A, ts(), rng.choice([b"", b"x", b"yy" * 20])))
                if rng.random() < 0.3: cells.append(Cell(Z, ts(), I32(rng.randint(-5, 5))))
                for u in [u for u in uuids if rng.random() < 0.15]:
                    cells.append(Cell(L_, ts(), rng.choice([b"e1", b"element-two", b""]), path=u) if rng.random() < 0.85 else Cell.tombstone(L_, ts(), NOW - rng.choice([5, 30 * 86400]), path=u))
                for k in sorted({rng.choice([b"a", b"b", b"c", b"dd", b"a-long-map-key" * 3]) for _ in range(rng.randint(0, 4))}):
                    r = rng.random()
                    if r < 0.7: cells.append(Cell(M, ts(), I32(rng.randint(0, 99)), path=k))
                    elif r < 0.85: cells.append(Cell.tombstone(M, ts(), NOW - rng.choice([5, 30 * 86400]), path=k))
                    else: cells.append(Cell(M, ts(), I32(1), 50, NOW + rng.choice([-30 * 86400, -100, 500]), path=k))
                for e in sorted({rng.randint(-4, 4) for _ in range(rng.randint(0, 3))}): cells.append(Cell(S_, ts(), b"", path=I64(e)))
                cd = {}
                if rng.random() < 0.25: cd[M] = (ts(), NOW - rng.choice([5, 30 * 86400]))
                if rng.random() < 0.1: cd[S_] = (ts(), NOW - rng.choice([5, 30 * 86400]))
                if rng.random() < 0.1: cd[L_] = (ts(), NOW - rng.choice([5, 30 * 86400]))
                dele = (ts(), NOW - rng.choice([5, 30 * 86400])) if rng.random() < 0.08 else None
                rts = ts() if (rng.random() < 0.7 or not (cells or cd or dele)) else NO_TS
                rows.append(Row((I32(ck),), cells, ts=rts, deletion=dele, complex_deletions=cd))
            parts.append(Partition(key, rows, (T0 + rng.randint(0, 400), NOW - 5) if rng.random() < 0.05 else None))
        import oracle_lib as O
        parts.sort(key=lambda p: (O.token(p.key), p.key))
        tables.append(Builder(SCX, (TIMESTAMP_EPOCH - t, DELETION_TIME_EPOCH, 0), column_index_size=cis).build(parts))
    return tables
