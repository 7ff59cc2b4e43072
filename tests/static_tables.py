"""Random tables with static columns for the oracle / K4 host build / GPU parity tests: static rows with live, expiring and
deleted cells, empty static rows, static-only partitions, partition deletions that shadow them, inputs whose header has fewer (or no) static
columns, and partitions wide enough for a promoted index (whose headerLength then covers the static row)."""
import random, struct
from sstable_builder import *

NOW = 1700000000
I32 = lambda v: struct.pack(">i", v)
I64 = lambda v: struct.pack(">q", v)
STATIC_ALL = [("s_blob", "BytesType"), ("s_long", "LongType"), ("s_text", "UTF8Type")]
REGULAR = [("a", "LongType"), ("b", "UTF8Type")]

def static_tables(seed, ntables=4, nkeys=80, cis=2048, wide=False):
    rng = random.Random(seed)
    keys = [b"k%04d" % i for i in range(nkeys)]
    tables = []
    for t in range(ntables):
        # table 1 has only a subset of the static columns, table 2 none at all (its partitions carry no static row)
        statics = STATIC_ALL if t % 3 == 0 else (STATIC_ALL[1:] if t % 3 == 1 else [])
        s = Schema(["Int32Type"], REGULAR, static_columns=statics)
        base_ts = 1000 + 10 * t
        def cell(ci, types):
            ts = base_ts + rng.randint(0, 30); kind = rng.random()
            tname = types[ci][1]
            val = I64(rng.getrandbits(40)) if tname.endswith("LongType") else bytes(rng.getrandbits(8) for _ in range(rng.choice([0, 1, 5, 40, 300])))
            if kind < 0.15: return Cell.tombstone(ci, ts, NOW - rng.choice([10, 10**6, 2 * 10**7]))
            if kind < 0.30: ttl = rng.choice([60, 86400 * 400]); return Cell(ci, ts, val, ttl, NOW - rng.choice([10**6, -10**6]) + ttl)
            return Cell(ci, ts, val)
        parts = []
        for k in keys:
            if rng.random() < 0.45: continue
            nrows = rng.choice([0, 0, 1, 3, 12]) if not wide else rng.choice([0, 40, 200])
            rows = []
            for ck in sorted(rng.sample(range(400), nrows)):
                cells = [cell(ci, s.columns) for ci in range(len(s.columns)) if rng.random() < 0.8]
                rows.append(Row((I32(ck),), cells, ts=base_ts + rng.randint(0, 30) if (rng.random() < 0.8 or not cells) else NO_TS))
            st = None
            if statics and rng.random() < 0.7:
                cells = [cell(ci, s.static_columns) for ci in range(len(s.static_columns)) if rng.random() < 0.7]
                if cells: st = Row((), cells)
            pdel = (base_ts + rng.randint(0, 30), NOW - rng.choice([10, 2 * 10**7])) if rng.random() < 0.15 else None
            if not rows and st is None and pdel is None: continue
            parts.append(Partition(k, rows, pdel, st))
        tab = Builder(s, (1000, NOW - 3 * 10**7, 60), column_index_size=cis).build(parts)
        tab.generation = t
        tables.append(tab)
    return tables
