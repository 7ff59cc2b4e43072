"""Counter columns through the oracle: the reference's own known answers for CounterContext.merge, the golden `oa` counter tables, and a randomised
comparison with an independent Python restatement (this file: a K-WAY, per-counter-id formulation — the oracle folds pairwise like the reference).
  CounterContext layout / merge rules / merge / compare     S/db/context/CounterContext.java:40-76, 296-447, 443-529
  ContextState (header walk, allocate, writeElement)        S/db/context/CounterContext.java:757-914
  hasLegacyShards                                           S/db/context/CounterContext.java:595-608
  Cells.reconcile -> resolveCounter                         S/db/rows/Cells.java:68-77, 121-162
  Cells.collectStats (updateHasLegacyCounterShards)         S/db/rows/Cells.java:44-50
  known answers                                             T/unit/org/apache/cassandra/db/context/CounterContextTest.java:261-380 (testMerge),
                                                            T/unit/org/apache/cassandra/db/CounterCellTest.java:127-186 (testReconcile)
  golden tables                                             T/data/legacy-sstables/oa/legacy_tables/legacy_oa_{simple,clust}_counter (LegacySSTableTest)"""
import ctypes as C, os, random, struct, copy, pytest
import oracle_lib as O
from sstable_builder import *
from cassandra_b200.io.sstable import SSTable
from cassandra_b200.db.compaction import CompactionTask, CompactionController, merged_encoding_stats
from test_oracle_complex_kats import raw_of, oracle_compact, reconcile, dt_sup, dt_deletes, live_sup, LIVE_DT, model_purge, I32

NOW = 1700000000
T0 = 1_600_000_000_000_000
G, L, R = "g", "l", "r"

def cid(n): return struct.pack(">qQ", 0, 0xC000000000000000 | n)            # CounterId.fromInt S/utils/CounterId.java:57-61
LOCAL_ID = struct.pack(">QQ", 0x0123456789AB11EE, 0x8000000000000077)        # any id above every fromInt(n) (the tests' getLocalId() sorts last there as well)

def ctx(shards):
    """shards: [(id16, clock, count, G|L|R)] in id order -> context bytes (ContextState.allocate + writeElement)"""
    elts = [(i - 32768 if k == G else i) for i, (_, _, _, k) in enumerate(shards) if k != R]
    out = struct.pack(">h", len(elts)) + b"".join(struct.pack(">h", e) for e in elts)
    return out + b"".join(i + struct.pack(">qq", cl, cn) for i, cl, cn, _ in shards)
def parse(b):
    (n,) = struct.unpack_from(">h", b, 0); n = abs(n)
    elts = struct.unpack_from(">%dh" % n, b, 2) if n else ()
    kinds = {}
    for e in elts: kinds[e + 32768 if e < 0 else e] = G if e < 0 else L
    body = b[2 + 2 * n:]; assert len(body) % 32 == 0
    return [(body[32 * i:32 * i + 16],) + struct.unpack_from(">qq", body, 32 * i + 16) + (kinds.get(i, R),) for i in range(len(body) // 32)]

def lib():
    Lb = O.lib()
    Lb.orc_counter_merge.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]; Lb.orc_counter_merge.restype = C.c_int
    Lb.orc_counter_has_legacy_shards.argtypes = [C.c_char_p, C.c_int]; Lb.orc_counter_has_legacy_shards.restype = C.c_int
    return Lb
def omerge(l, r):
    buf = C.create_string_buffer(len(l) + len(r) + 16); w = C.c_int(-1)
    n = lib().orc_counter_merge(l, len(l), r, len(r), buf, len(buf), C.byref(w))
    assert n >= 0
    return buf.raw[:n], w.value

# ---- the reference's known answers (CounterContextTest.testMerge) --------------------------------------------------------------------------
def test_reference_merge_local_counts_add_remote_counts_reconcile():
    """:264-301: local shards of the same id add clocks and counts, remote shards keep the larger clock"""
    left = ctx([(cid(1), 1, 1, R), (cid(2), 2, 2, R), (cid(4), 6, 3, R), (LOCAL_ID, 7, 3, L)])
    right = ctx([(cid(4), 4, 4, R), (cid(5), 5, 5, R), (LOCAL_ID, 2, 9, L)])
    m, which = omerge(left, right)
    assert which == 2 and len(m) == 4 + 5 * 32                                          # hd = 4: one header element
    assert parse(m) == [(cid(1), 1, 1, R), (cid(2), 2, 2, R), (cid(4), 6, 3, R), (cid(5), 5, 5, R), (LOCAL_ID, 9, 12, L)]

def test_reference_merge_of_global_contexts():
    """:303-338: the union, the larger clock for the shared id; 5 header elements; total 18"""
    left = ctx([(cid(1), 1, 1, G), (cid(2), 2, 2, G), (cid(3), 3, 3, G)])
    right = ctx([(cid(3), 6, 6, G), (cid(4), 4, 4, G), (cid(5), 5, 5, G)])
    m, _ = omerge(left, right)
    assert len(m) == 2 + 5 * 2 + 5 * 32 and struct.unpack_from(">h", m, 0)[0] == 5
    assert parse(m) == [(cid(i), c, c, G) for i, c in ((1, 1), (2, 2), (3, 6), (4, 4), (5, 5))]
    assert sum(s[2] for s in parse(m)) == 18

def test_reference_merge_invalid_global_shards_pick_the_larger_count():
    """:340-357: equal clocks, different counts"""
    m, which = omerge(ctx([(cid(1), 10, 20, G)]), ctx([(cid(1), 10, 30, G)]))
    assert parse(m) == [(cid(1), 10, 30, G)] and which == 1                             # the right context is a superset: returned as it is

def test_reference_merge_global_dominates_local_and_remote():
    """:359-379: global shards win even with lower clock and value"""
    left = ctx([(cid(1), 1, 1, G), (cid(2), 1, 1, G)])
    right = ctx([(cid(1), 100, 100, L), (cid(2), 100, 100, R)])
    m, which = omerge(left, right)
    assert parse(m) == [(cid(1), 1, 1, G), (cid(2), 1, 1, G)] and which == 0
    m2, which2 = omerge(right, left)
    assert m2 == m and which2 == 1

def test_has_legacy_shards():
    hl = lambda b: lib().orc_counter_has_legacy_shards(b, len(b))
    assert hl(ctx([(cid(1), 1, 1, G), (cid(2), 1, 1, G)])) == 0
    assert hl(ctx([(cid(1), 1, 1, G), (cid(2), 1, 1, R)])) == 1 and hl(ctx([(cid(1), 1, 1, L)])) == 1 and hl(ctx([])) == 0

# ---- the model: per counter id over ALL contexts at once ------------------------------------------------------------------------------------
def model_ctx_merge(contexts):
    """globals dominate (largest (clock, count)); else the locals add up; else the remotes' largest (clock, count)  [CounterContext.java:66-73]"""
    by_id = {}
    for c in contexts:
        for i, cl, cn, k in parse(c): by_id.setdefault(i, []).append((cl, cn, k))
    out = []
    wrap = lambda v: (v + (1 << 63)) % (1 << 64) - (1 << 63)
    for i in sorted(by_id):
        v = by_id[i]
        gs = [x for x in v if x[2] == G]; ls = [x for x in v if x[2] == L]
        if gs: cl, cn, _ = max(gs); out.append((i, cl, cn, G))
        elif ls: out.append((i, wrap(sum(x[0] for x in ls)), wrap(sum(x[1] for x in ls)), L))
        else: cl, cn, _ = max(v); out.append((i, cl, cn, R))
    return ctx(out)

def model_resolve_counter(cells):
    """Cells.resolveCounter folded over the versions' cells (source order)"""
    tombs = [c for c in cells if c.ldt != NO_DELETION_TIME and c.ttl == 0]
    if tombs:                                                                            # a tombstone beats every counter cell; tombstones among themselves: resolveRegular
        m = tombs[0]
        for c in tombs[1:]: m = reconcile(m, c)
        return m
    empties = [c for c in cells if not c.value]
    if empties:                                                                          # :142-149: an empty value beats a context; among empties the larger timestamp, the later one on a tie
        m = empties[0]
        for c in empties[1:]: m = m if m.ts > c.ts else c
        return m
    return Cell(cells[0].col, max(c.ts for c in cells), model_ctx_merge([c.value for c in cells]))

def model_merge_row(versions, active, ncols):
    if len(versions) == 1 and active == LIVE_DT: return versions[0]
    info, dele = (NO_TS, 0, NO_DELETION_TIME), LIVE_DT
    for v in versions:
        vi = (v.ts, v.ttl, v.ldt)
        if live_sup(vi, info): info = vi
        vd = v.deletion or LIVE_DT
        if dt_sup(vd, dele): dele = vd
    if dt_sup(dele, active): active = dele
    else: dele = LIVE_DT
    if dt_deletes(active, info[0]): info = (NO_TS, 0, NO_DELETION_TIME)
    out = Row(versions[0].ck, [], info[0], info[1], info[2], None if dele == LIVE_DT else dele)
    for col in range(ncols):
        cs = [c for v in versions for c in v.cells if c.col == col and not dt_deletes(active, c.ts)]      # skipped BEFORE the merge (Row.java:838-849): the merged timestamp is the max
        if cs: out.cells.append(model_resolve_counter(cs) if len(cs) > 1 else cs[0])
    if out.ts == NO_TS and out.deletion is None and not out.cells: return None
    return out

def model_compact(schema, tables_parts, now=NOW, gc_grace=864000):
    keys = {}
    for parts in tables_parts:
        for p in parts: keys.setdefault(p.key, []).append(p)
    out = []
    for key in sorted(keys, key=lambda k: (O.token(k), k)):
        ps = keys[key]; pdel = LIVE_DT
        for p in ps:
            if p.deletion is not None and not dt_sup(pdel, p.deletion): pdel = p.deletion
        cks = sorted({u.ck for p in ps for u in p.unfiltereds}, key=lambda ck: struct.unpack(">i", ck[0])[0])
        rows = []
        for ck in cks:
            vs = [u for p in ps for u in p.unfiltereds if u.ck == ck]
            m = model_merge_row(vs, pdel, len(schema.columns)) if len(ps) > 1 else vs[0]
            if m is not None: m = model_purge(copy.deepcopy(m), now, now - gc_grace)
            if m is not None: rows.append(m)
        out_pdel = pdel if pdel != LIVE_DT and not (pdel[1] < now - gc_grace) else None
        if rows or out_pdel is not None: out.append(Partition(key, rows, out_pdel))
    return out

SC = Schema(["Int32Type"], [("a", "CounterColumnType"), ("b", "CounterColumnType")])
def check(tables_parts, now=NOW, gc_grace=864000, schema=SC):
    tabs = [Builder(schema).build(parts) for parts in tables_parts]
    got = oracle_compact(tabs, now, gc_grace)
    want_parts = model_compact(schema, tables_parts, now, gc_grace)
    want = raw_of(Builder(schema, merged_encoding_stats(tabs)).build(want_parts)) if want_parts else b""
    assert got == want
    return got

# ---- rows ---------------------------------------------------------------------------------------------------------------------------------
def test_counter_cells_of_one_row_merge_shard_by_shard():
    c1 = ctx([(cid(1), 3, 30, G), (cid(2), 1, 5, G)]); c2 = ctx([(cid(2), 2, 9, G), (cid(3), 1, 1, G)]); c3 = ctx([(cid(1), 2, 20, G), (cid(4), 7, 7, G)])
    t = [[Partition(b"k", [Row((I32(1),), [Cell(0, T0 + i, c)])])] for i, c in enumerate((c1, c2, c3))]
    raw = check(t)
    assert ctx([(cid(1), 3, 30, G), (cid(2), 2, 9, G), (cid(3), 1, 1, G), (cid(4), 7, 7, G)]) in raw      # 4 shards, timestamp T0 + 2

def test_superset_context_is_kept_but_the_timestamp_is_the_largest():
    big = ctx([(cid(1), 5, 50, G), (cid(2), 5, 50, G)]); small = ctx([(cid(1), 4, 40, G)])
    t = [[Partition(b"k", [Row((I32(1),), [Cell(0, T0, big)])])], [Partition(b"k", [Row((I32(1),), [Cell(0, T0 + 9, small)])])]]
    raw = check(t)
    assert big in raw and small not in raw

def test_tombstone_beats_counter_cells_whatever_the_timestamps():
    c1 = ctx([(cid(1), 3, 30, G)])
    t = [[Partition(b"k", [Row((I32(1),), [Cell(0, T0 + 100, c1)])])], [Partition(b"k", [Row((I32(1),), [Cell.tombstone(0, T0, NOW - 5)])])],
         [Partition(b"k", [Row((I32(1),), [Cell(0, T0 + 200, c1), Cell(1, T0, c1)])])]]
    raw = check(t)
    assert raw.count(c1) == 1                                                            # column a is the tombstone, column b passes
    check(t, gc_grace=1)                                                                 # ... and the purgeable tombstone disappears

def test_row_and_partition_deletions_skip_cells_before_the_merge():
    c1 = ctx([(cid(1), 3, 30, G)]); c2 = ctx([(cid(2), 1, 1, G)]); c3 = ctx([(cid(3), 1, 1, G)])
    t = [[Partition(b"k", [Row((I32(1),), [Cell(0, T0 + 5, c1)])])],
         [Partition(b"k", [Row((I32(1),), [Cell(0, T0 + 50, c2)], deletion=(T0 + 10, NOW - 5))])],
         [Partition(b"k", [Row((I32(1),), [Cell(0, T0 + 60, c3)]), Row((I32(2),), [Cell(0, T0 + 7, c3)])], deletion=(T0 + 8, NOW - 5))]]
    raw = check(t)
    assert c1 not in raw and ctx([(cid(2), 1, 1, G), (cid(3), 1, 1, G)]) in raw          # the shard of the deleted cell is NOT part of the merged context

def test_legacy_local_and_remote_shards():
    l1 = ctx([(cid(1), 2, 10, R), (cid(7), 3, 4, L)]); l2 = ctx([(cid(1), 5, 1, R), (cid(7), 1, 1, L)]); g = ctx([(cid(7), 1, 100, G)])
    raw = check([[Partition(b"k", [Row((I32(1),), [Cell(0, T0, l1)])])], [Partition(b"k", [Row((I32(1),), [Cell(0, T0, l2)])])]])
    assert ctx([(cid(1), 5, 1, R), (cid(7), 4, 5, L)]) in raw
    raw = check([[Partition(b"k", [Row((I32(1),), [Cell(0, T0, l1)])])], [Partition(b"k", [Row((I32(1),), [Cell(0, T0, g)])])], [Partition(b"k", [Row((I32(1),), [Cell(0, T0, l2)])])]])
    assert ctx([(cid(1), 5, 1, R), (cid(7), 1, 100, G)]) in raw

def test_statistics_flag_for_legacy_shards():
    def stats_of(value):
        tabs = [Builder(SC).build([Partition(b"k", [Row((I32(1),), [Cell(0, T0, value)])])])]
        for g_, t in enumerate(tabs): t.generation = g_
        r = CompactionTask(tabs, CompactionController(NOW), bloom=(1, 1)).execute(O.OracleEngine())
        return r.outputs[0].stats
    assert stats_of(ctx([(cid(1), 1, 1, G)]))["has_legacy_counter_shards"] == 0
    assert stats_of(ctx([(cid(1), 1, 1, G), (cid(2), 1, 1, R)]))["has_legacy_counter_shards"] == 1
    assert stats_of(ctx([(cid(1), 1, 1, L)]))["has_legacy_counter_shards"] == 1

def test_malformed_context_is_corruption_not_garbage():
    from cassandra_b200 import native
    bad = ctx([(cid(1), 1, 1, G)])[:-3]
    good = ctx([(cid(1), 1, 1, G)])
    tabs = [Builder(SC).build([Partition(b"k", [Row((I32(1),), [Cell(0, T0, v)])])]) for v in (bad, good)]
    for g_, t in enumerate(tabs): t.generation = g_
    with pytest.raises(native.CorruptSSTableError):
        CompactionTask(tabs, CompactionController(NOW)).execute(O.OracleEngine())

def random_context(rng, ids):
    shards = []
    for i in sorted(rng.sample(ids, rng.randint(0, len(ids)))):
        shards.append((cid(i), rng.choice([1, 2, 2, 3, 1 << 40, -5]), rng.choice([0, 1, 7, -3, 1 << 50, (1 << 63) - 1]), rng.choice([G, G, G, L, R])))
    return ctx(shards)

def counter_tables(rng, ntables, nkeys=4, nck=4):
    tables = []
    for _ in range(ntables):
        parts = []
        for k in range(nkeys):
            if rng.random() < 0.25: continue
            rows = []
            for ck in range(nck):
                if rng.random() < 0.3: continue
                cells = []
                for col in (0, 1):
                    x = rng.random()
                    if x < 0.2: continue
                    ts = T0 + rng.randint(0, 40)
                    if x < 0.3: cells.append(Cell.tombstone(col, ts, NOW - rng.choice([5, 5, 30 * 86400])))
                    elif x < 0.36: cells.append(Cell(col, ts, b""))                           # (an empty value: beats contexts, Cells.java:142-149)
                    else: cells.append(Cell(col, ts, random_context(rng, [1, 2, 3, 4, 5])))
                dele = (T0 + rng.randint(0, 40), NOW - rng.choice([5, 30 * 86400])) if rng.random() < 0.15 else None
                if cells or dele: rows.append(Row((I32(ck),), cells, deletion=dele))
            pdel = (T0 + rng.randint(0, 30), NOW - rng.choice([5, 30 * 86400])) if rng.random() < 0.15 else None
            if rows or pdel: parts.append(Partition(b"key%d" % k, rows, pdel))
        tables.append(parts)
    return tables

def test_randomised_against_the_model():
    rng = random.Random(0xC0FFEE)
    for it in range(80):
        tables = counter_tables(rng, rng.randint(1, 5))
        if not any(tables): continue
        tables = [t for t in tables if t]
        check(tables, gc_grace=rng.choice([864000, 1, 10 ** 9]))

# ---- golden tables written by a real Cassandra release ------------------------------------------------------------------------------------
def _golden(golden_dir, name): return os.path.join(golden_dir, "oa", "legacy_tables", name, "oa-1-big-")

@pytest.mark.parametrize("name", ["legacy_oa_simple_counter", "legacy_oa_clust_counter"])
def test_identity_compaction_reproduces_the_golden_counter_tables(golden_dir, name):
    base = _golden(golden_dir, name)
    s = SSTable.open(base)
    assert s.regular_columns[0][1].endswith("CounterColumnType")
    import struct as _st
    hc, words = _st.unpack_from(">ii", open(base + "Filter.db", "rb").read(), 0)
    r = CompactionTask([s], CompactionController(NOW), column_index_size=4096, bloom=(hc, words)).execute(O.OracleEngine())
    comp = r.outputs[0].components()
    for c in ("Data.db", "Index.db", "CompressionInfo.db", "Digest.crc32", "Filter.db", "Summary.db"):
        assert comp[c] == open(base + c, "rb").read(), c
    assert r.outputs[0].stats["has_legacy_counter_shards"] == 0                          # written by 5.0: global shards only

@pytest.mark.parametrize("name", ["legacy_oa_simple_counter", "legacy_oa_clust_counter"])
def test_golden_counter_table_merged_with_itself(golden_dir, name):
    """every shard meets its equal: each context is a superset of the other and the left one is returned — the file reproduces itself"""
    base = _golden(golden_dir, name)
    a, b, c = SSTable.open(base, 1), SSTable.open(base, 2), SSTable.open(base, 3)
    r = CompactionTask([a, b, c], CompactionController(NOW), column_index_size=4096).execute(O.OracleEngine())
    assert r.outputs[0].components()["Data.db"] == open(base + "Data.db", "rb").read()
    assert r.outputs[0].components()["Index.db"] == open(base + "Index.db", "rb").read()

# ---- static counter columns ---------------------------------------------------------------------------------------------------------------
SCS = Schema(["Int32Type"], [("a", "CounterColumnType")], static_columns=[("s", "CounterColumnType"), ("t", "CounterColumnType")])

def model_compact_static(schema, tables_parts, now=NOW, gc_grace=864000):
    """model_compact + the static row: merged eagerly with the partition deletion as the active deletion, whatever the fan-in
    (mergeStaticRows S/db/rows/UnfilteredRowIterators.java:484-505), then purged like a row"""
    keys = {}
    for parts in tables_parts:
        for p in parts: keys.setdefault(p.key, []).append(p)
    out = []
    for key in sorted(keys, key=lambda k: (O.token(k), k)):
        ps = keys[key]; pdel = LIVE_DT
        for p in ps:
            if p.deletion is not None and not dt_sup(pdel, p.deletion): pdel = p.deletion
        svs = [p.static for p in ps if p.static is not None]
        st = None
        if svs:
            st = model_merge_row(svs, pdel, len(schema.static_columns)) if (len(ps) > 1 or pdel != LIVE_DT) else svs[0]
            if st is not None: st = model_purge(copy.deepcopy(st), now, now - gc_grace)
        cks = sorted({u.ck for p in ps for u in p.unfiltereds}, key=lambda ck: struct.unpack(">i", ck[0])[0])
        rows = []
        for ck in cks:
            vs = [u for p in ps for u in p.unfiltereds if u.ck == ck]
            m = model_merge_row(vs, pdel, len(schema.columns)) if len(ps) > 1 else vs[0]
            if m is not None: m = model_purge(copy.deepcopy(m), now, now - gc_grace)
            if m is not None: rows.append(m)
        out_pdel = pdel if pdel != LIVE_DT and not (pdel[1] < now - gc_grace) else None
        if rows or out_pdel is not None or st is not None: out.append(Partition(key, rows, out_pdel, static=st))
    return out

def check_static(tables_parts, now=NOW, gc_grace=864000):
    tabs = [Builder(SCS).build(parts) for parts in tables_parts]
    got = oracle_compact(tabs, now, gc_grace)
    want_parts = model_compact_static(SCS, tables_parts, now, gc_grace)
    want = raw_of(Builder(SCS, merged_encoding_stats(tabs)).build(want_parts)) if want_parts else b""
    assert got == want
    return got

def test_static_counter_cells_merge_like_regular_ones():
    c1 = ctx([(cid(1), 3, 30, G)]); c2 = ctx([(cid(2), 1, 1, G)]); c3 = ctx([(cid(1), 9, 90, G)])
    t = [[Partition(b"k", [Row((I32(1),), [Cell(0, T0, c1)])], static=Row((), [Cell(0, T0 + 1, c1), Cell(1, T0, c2)]))],
         [Partition(b"k", [], static=Row((), [Cell(0, T0 + 2, c2)]))],
         [Partition(b"k", [Row((I32(1),), [Cell(0, T0 + 3, c2)])], static=Row((), [Cell(0, T0, c3), Cell.tombstone(1, T0 - 5, NOW - 5)]))]]
    raw = check_static(t)
    assert ctx([(cid(1), 9, 90, G), (cid(2), 1, 1, G)]) in raw                           # static s: three contexts merged; static t: the tombstone wins
    check_static([t[0], [Partition(b"k", [], (T0 + 1, NOW - 5), static=Row((), [Cell(0, T0 + 7, c2)]))]])      # partition deletion skips the older static cells before the merge

def test_static_counters_randomised_against_the_model():
    rng = random.Random(0x57A71C)
    for it in range(60):
        tables = []
        for _ in range(rng.randint(1, 4)):
            parts = []
            for k in range(4):
                if rng.random() < 0.3: continue
                rows = [Row((I32(ck),), [Cell(0, T0 + rng.randint(0, 40), random_context(rng, [1, 2, 3]))]) for ck in range(3) if rng.random() < 0.5]
                scells = []
                for col in (0, 1):
                    x = rng.random(); ts = T0 + rng.randint(0, 40)
                    if x < 0.3: continue
                    if x < 0.4: scells.append(Cell.tombstone(col, ts, NOW - rng.choice([5, 30 * 86400])))
                    else: scells.append(Cell(col, ts, random_context(rng, [1, 2, 3, 4])))
                st = Row((), scells) if scells else None
                pdel = (T0 + rng.randint(0, 30), NOW - rng.choice([5, 30 * 86400])) if rng.random() < 0.2 else None
                if rows or st or pdel: parts.append(Partition(b"key%d" % k, rows, pdel, static=st))
            if parts: tables.append(parts)
        if tables: check_static(tables, gc_grace=rng.choice([864000, 1, 10 ** 9]))

# ---- the reference's known answers for Cells.reconcile on counter cells (T/unit/org/apache/cassandra/db/CounterCellTest.java:127-186, testReconcile) --------
def _one_cell_tables(cells):
    return [[Partition(b"k", [Row((I32(1),), [c])])] for c in cells]
def _expect_cell(tables_parts, cell, gc_grace=NOW + 10):
    """the compaction of the one-cell tables holds exactly `cell` (None: nothing is left); gc_before < 0: the test's tiny deletion times are not purgeable"""
    tabs = [Builder(SC).build(p) for p in tables_parts]
    got = oracle_compact(tabs, NOW, gc_grace)
    want = raw_of(Builder(SC, merged_encoding_stats(tabs)).build([Partition(b"k", [Row((I32(1),), [cell])])])) if cell is not None else b""
    assert got == want

def test_reference_counter_cell_reconcile():
    local = lambda count: ctx([(LOCAL_ID, 1, count, L)])                                 # CounterContext.createLocal(count): one local shard, clock 1 (:139-144)
    dead = lambda ts, ldt: Cell.tombstone(0, ts, ldt)
    _expect_cell(_one_cell_tables([dead(2, 5), dead(2, 10)]), dead(2, 10))              # :137-140 both deleted, same ts: the later deletion time
    _expect_cell(_one_cell_tables([dead(2, 5), Cell(0, 10, local(1))]), dead(2, 5))     # :142-144 "diff ts": the tombstone, although older
    _expect_cell(_one_cell_tables([dead(6, 6), Cell(0, 5, local(1))]), dead(6, 6))      # :146-149
    _expect_cell(_one_cell_tables([dead(1, 1), Cell(0, 5, local(1))]), dead(1, 1))      # :151-154
    _expect_cell(_one_cell_tables([dead(8, 8), Cell(0, 8, local(1))]), dead(8, 8))      # :156-159
    live = [Cell(0, 2, local(1)), Cell(0, 5, local(3))]
    _expect_cell(_one_cell_tables(live), Cell(0, 5, ctx([(LOCAL_ID, 2, 4, L)])))        # :161-166 live + live: total 4, timestamp 5
    live.append(Cell(0, 4, local(10)))
    _expect_cell(_one_cell_tables(live), Cell(0, 5, ctx([(LOCAL_ID, 3, 14, L)])))       # :168-172 add, the timestamp stays
    live.append(Cell(0, 7, local(3)))
    _expect_cell(_one_cell_tables(live), Cell(0, 7, ctx([(LOCAL_ID, 4, 17, L)])))       # :174-178 add with a newer timestamp
    _expect_cell(_one_cell_tables(live + [dead(8, 8)]), dead(8, 8))                     # :183-185 ... and a tombstone ends it
