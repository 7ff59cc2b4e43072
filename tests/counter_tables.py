"""Randomised counter tables — two counter columns, contexts of global / local / remote shards, tombstones, empty values, row and partition deletions —
for the K4 host-build and GPU parity tests (the oracle they are compared with is pinned by tests/test_oracle_counter_kats.py)."""
import random, struct
from sstable_builder import *

NOW = 1700000000
I32 = lambda v: struct.pack(">i", v)
T0 = 1_600_000_000_000_000
SCTR = Schema(["Int32Type"], [("a", "CounterColumnType"), ("b", "CounterColumnType")])
G, L, R = "g", "l", "r"

def cid(n): return struct.pack(">qQ", 0, 0xC000000000000000 | n)
def ctx(shards):
    """[(id16, clock, count, G|L|R)] in id order -> counter context (S/db/context/CounterContext.java:40-76)"""
    elts = [(i - 32768 if k == G else i) for i, (_, _, _, k) in enumerate(shards) if k != R]
    return struct.pack(">h", len(elts)) + b"".join(struct.pack(">h", e) for e in elts) + b"".join(i + struct.pack(">qq", cl, cn) for i, cl, cn, _ in shards)

def random_context(rng, ids, legacy=True):
    kinds = [G, G, G, L, R] if legacy else [G]
    return ctx([(cid(i), rng.choice([1, 2, 2, 3, 1 << 40, -5]), rng.choice([0, 1, 7, -3, 1 << 50, (1 << 63) - 1]), rng.choice(kinds))
                for i in sorted(rng.sample(ids, rng.randint(0, len(ids))))])

SCTR_STATIC = Schema(["Int32Type"], [("a", "CounterColumnType"), ("b", "CounterColumnType")], static_columns=[("s", "CounterColumnType"), ("t", "CounterColumnType")])

def counter_tables(seed, ntables=4, nkeys=60, cis=65536, big=False, legacy=True, static=False):
    """static: the table also has two static counter columns (a static row per partition, merged with the partition deletion in force)"""
    rng = random.Random(seed)
    schema = SCTR_STATIC if static else SCTR
    keys = sorted({b"key%05d" % rng.randint(0, 10 ** 5) for _ in range(nkeys)})
    ids = list(range(1, 9 if not big else 40))
    tables = []
    for t in range(ntables):
        parts = []
        for key in keys:
            if rng.random() < 0.35: continue
            rows = []
            for ck in sorted({rng.randint(0, 5 if not big else 300) for _ in range(rng.randint(1, 4) if not big else 120)}):
                cells = []
                for col in (0, 1):
                    x = rng.random(); ts = T0 + rng.randint(0, 60)
                    if x < 0.2: continue
                    if x < 0.28: cells.append(Cell.tombstone(col, ts, NOW - rng.choice([5, 5, 30 * 86400])))
                    elif x < 0.32: cells.append(Cell(col, ts, b""))
                    else: cells.append(Cell(col, ts, random_context(rng, ids, legacy)))
                dele = (T0 + rng.randint(0, 60), NOW - rng.choice([5, 30 * 86400])) if rng.random() < 0.12 else None
                if cells or dele: rows.append(Row((I32(ck),), cells, deletion=dele))
            pdel = (T0 + rng.randint(0, 40), NOW - rng.choice([5, 30 * 86400])) if rng.random() < 0.08 else None
            st = None
            if static and rng.random() < 0.7:
                scells = []
                for col in (0, 1):
                    x = rng.random(); ts = T0 + rng.randint(0, 60)
                    if x < 0.25: continue
                    if x < 0.33: scells.append(Cell.tombstone(col, ts, NOW - rng.choice([5, 5, 30 * 86400])))
                    else: scells.append(Cell(col, ts, random_context(rng, ids, legacy)))
                if scells: st = Row((), scells)
            if rows or pdel or st: parts.append(Partition(key, rows, pdel, static=st))
        tables.append(Builder(schema, column_index_size=cis).build(parts))
    for g, tb in enumerate(tables): tb.generation = g
    return tables
