"""Known-answer and property tests of the oracle's merge / reconcile / purge semantics, ported from the reference's own tests:
  T/db/CellTest.java:268-350 (reconcile + purge / TTL->tombstone conversion KATs)
  T/db/rows/RowsTest.java:461-511 (row merge: deletion superseding liveness)
  T/db/rows/UnfilteredRowIteratorsMergeTest.java:72-104 (randomised range-tombstone merges checked by an independent model)
  T/db/compaction/CompactionsPurgeTest.java (gcBefore / max-timestamp purge rule)
The inputs are written by the independent Python writer in sstable_builder.py, compacted by the oracle, decoded back."""
import random, struct, pytest
import oracle_lib as O
from sstable_builder import *
from cassandra_b200.db.compaction import CompactionTask, CompactionController

NOW = 1700000000
I32 = lambda v: struct.pack(">i", v)
S1 = Schema(["Int32Type"], [("val", "UTF8Type")])

def compact(tables, now=NOW, gc_grace=864000, overlap_min_ts=None, schema=S1, cis=65536):
    for g, t in enumerate(tables): t.generation = g
    r = CompactionTask(tables, CompactionController(now, gc_grace, overlap_min_ts), column_index_size=cis).execute(O.OracleEngine())
    assert len(r.outputs) == 1
    o = r.outputs[0]
    raw = b"".join(O.chunk_decompress(O.COMP_LZ4, o.data[a:(b if b else len(o.data)) - 4], 16384)
                   for a, b in zip(o.compression.chunk_offsets, o.compression.chunk_offsets[1:] + [None]))
    from cassandra_b200.db.compaction import merged_encoding_stats
    return decode_stream(schema, raw, merged_encoding_stats(tables)), r

def one_cell_table(cell, stats=(0, 0, 0)):
    return Builder(S1, stats).build([Partition(b"k", [Row((I32(1),), [cell])])])

def merged_cell(a, b, **kw):
    parts, _ = compact([one_cell_table(a), one_cell_table(b)], **kw)
    if not parts: return None
    cells = parts[0].unfiltereds[0].cells
    return cells[0] if cells else None

def cell_tuple(c): return None if c is None else (c.ts, c.ttl, c.ldt, c.value)

def test_reconcile_kats():
    # CellTest.testExpiringCellReconile (:268-290): (ts, ttl, ldt, value) pairs and which one wins
    far = NOW + 10**6
    live = lambda ts, v: Cell(0, ts, v)
    exp = lambda ts, v, ttl, ldt: Cell(0, ts, v, ttl, ldt)
    tomb = lambda ts, ldt: Cell.tombstone(0, ts, ldt)
    cases = [
        (live(2, b"a"), live(1, b"b"), 0),                         # higher timestamp wins
        (exp(1, b"a", 100, far), live(1, b"z"), 0),                # same ts: expiring beats live (CASSANDRA-14592)
        (tomb(1, far), live(1, b"z"), 0),                          # same ts: tombstone beats live
        (tomb(1, far), exp(1, b"a", 100, far + 5), 0),             # same ts: pure tombstone beats expiring
        (exp(1, b"a", 100, far + 2), exp(1, b"b", 100, far + 1), 0),   # both expiring: higher localDeletionTime
        (tomb(1, far + 2), tomb(1, far + 1), 0),
        (live(1, b"b"), live(1, b"a"), 0),                         # all equal: greater value
        (live(1, b"ab"), live(1, b"a"), 0),
        (exp(1, b"b", 100, far), exp(1, b"a", 100, far), 0),
    ]
    for a, b, win in cases:
        w = (a, b)[win]
        assert cell_tuple(merged_cell(a, b, gc_grace=0, now=NOW)) == cell_tuple(w)
        assert cell_tuple(merged_cell(b, a, gc_grace=0, now=NOW)) == cell_tuple(w)      # commutative (Cells.java:58)

def test_cell_purge_kats():
    # CellTest :305-350 with now = 100 style values scaled to seconds-since-epoch ranges the format can carry
    now = 1000
    t = lambda c, gcb, nowsec: cell_tuple(merged_cell(c, Cell(0, -5, b"old"), now=nowsec, gc_grace=nowsec - gcb))
    tomb = Cell.tombstone(0, now, now)
    assert t(tomb, now - 1, now + 1) == cell_tuple(tomb)                      # testNonPurgableTombstone
    assert t(tomb, now + 1, now + 1) is None                                  # testPurgeableTombstone (cell, hence row, vanish)
    live_exp = Cell(0, now, b"a", 10, now + 10)
    assert t(live_exp, now, now + 1) == cell_tuple(live_exp)                  # testLiveExpiringCell
    assert t(live_exp, now, now + 11) == (now, 0, now, b"")                   # testExpiredTombstoneConversion: ldt - ttl, value dropped
    assert t(live_exp, now + 1, now + 11) is None                             # testPurgeableExpiringCell

def test_overlapping_sstable_min_timestamp_blocks_purge():
    # CompactionsPurgeTest: a tombstone newer than data in a non-compacting sstable must survive
    tomb = Cell.tombstone(0, 500, 1000)
    assert cell_tuple(merged_cell(tomb, Cell(0, 1, b"x"), now=NOW, overlap_min_ts=400)) == cell_tuple(tomb)
    assert merged_cell(tomb, Cell(0, 1, b"x"), now=NOW, overlap_min_ts=501) is None
    assert merged_cell(tomb, Cell(0, 1, b"x"), now=NOW) is None

def test_row_deletion_and_liveness_merge():
    # RowsTest.merge / mergeRowDeletionSupercedesLiveness (:461-511)
    b = Builder(S1, (0, 0, 0))
    k = (I32(7),)
    t1 = b.build([Partition(b"k", [Row(k, [Cell(0, 10, b"a")], ts=10)])])
    t2 = b.build([Partition(b"k", [Row(k, [], deletion=(20, NOW))])])           # row deletion newer than everything
    parts, _ = compact([t1, t2], gc_grace=10**9)
    r = parts[0].unfiltereds[0]
    assert r.deletion == (20, NOW) and r.ts == NO_TS and r.cells == []
    t3 = b.build([Partition(b"k", [Row(k, [Cell(0, 30, b"c")], ts=30)])])        # newer write survives the deletion
    parts, _ = compact([t1, t2, t3], gc_grace=10**9)
    r = parts[0].unfiltereds[0]
    assert r.deletion == (20, NOW) and r.ts == 30 and [cell_tuple(c) for c in r.cells] == [(30, 0, NO_DELETION_TIME, b"c")]
    # liveness: same timestamp, expiring supersedes non-expiring (LivenessInfo.supersedes :216-225)
    t4 = b.build([Partition(b"k", [Row(k, [], ts=10, ttl=50, ldt=NOW + 50)])])
    parts, _ = compact([t1, t4], gc_grace=10**9)
    assert (parts[0].unfiltereds[0].ts, parts[0].unfiltereds[0].ttl) == (10, 50)

def test_partition_deletion_shadows_and_purges():
    b = Builder(S1, (0, 0, 0))
    t1 = b.build([Partition(b"a", [Row((I32(1),), [Cell(0, 10, b"x")], ts=10), Row((I32(2),), [Cell(0, 40, b"y")], ts=40)])])
    t2 = b.build([Partition(b"a", [], deletion=(20, NOW - 100))])
    parts, r = compact([t1, t2], gc_grace=10**9)            # not purgeable: deletion kept, shadowed row dropped
    assert parts[0].deletion == (20, NOW - 100) and [u.ck for u in parts[0].unfiltereds] == [(I32(2),)]
    parts, r = compact([t1, t2], gc_grace=10)               # purgeable: deletion gone but it still shadowed row 1 during the merge
    assert parts[0].deletion is None and [u.ck for u in parts[0].unfiltereds] == [(I32(2),)]
    t3 = b.build([Partition(b"b", [], deletion=(20, NOW - 100))])
    parts, r = compact([t3], gc_grace=10)                   # partition with only a purgeable deletion disappears entirely
    assert parts == [] and r.outputs[0].data == b"" and r.outputs[0].index == b""
    # single source: TrivialOneToOne passes rows through even when the partition deletion shadows them (no Row.Merger call)
    t4 = b.build([Partition(b"a", [Row((I32(1),), [Cell(0, 10, b"x")], ts=10)], deletion=(20, NOW - 100))])
    parts, r = compact([t4], gc_grace=10**9)
    assert parts[0].deletion == (20, NOW - 100) and len(parts[0].unfiltereds) == 1

# ---- range tombstones: DSL of T/db/rows/UnfilteredRowsGenerator.java:187-216 + independent semantic model ------------------
import re
def parse_dsl(s, default_liveness=100):
    """'5<=[140] 10[150] [140]<20': open incl at 5 del 140, row 10 ts 150, close excl at 20. Adjacent close/open at the same
    position are joined into a boundary (attachBoundaries)."""
    out = []
    for tok in s.split():
        m = re.fullmatch(r"(\d+)<(=)?\[(\d+)\]", tok)
        if m: out.append(Marker(K_INCL_START if m.group(2) else K_EXCL_START, (I32(int(m.group(1))),), None, (int(m.group(3)), int(m.group(3))))); continue
        m = re.fullmatch(r"\[(\d+)\]<(=)?(\d+)", tok)
        if m: out.append(Marker(K_INCL_END if m.group(2) else K_EXCL_END, (I32(int(m.group(3))),), (int(m.group(1)), int(m.group(1))), None)); continue
        m = re.fullmatch(r"(\d+)(\[(\d+)(?:D(\d+))?\])?", tok)
        live = int(m.group(3)) if m.group(3) else default_liveness
        d = (int(m.group(4)), int(m.group(4))) if m.group(4) else None
        out.append(Row((I32(int(m.group(1))),), [], ts=live, deletion=d))
    joined = []
    for u in out:
        p = joined[-1] if joined else None
        if isinstance(u, Marker) and isinstance(p, Marker) and p.open is None and u.close is None and p.ck == u.ck and \
           ((p.kind == K_EXCL_END and u.kind == K_INCL_START) or (p.kind == K_INCL_END and u.kind == K_EXCL_START)):
            joined[-1] = Marker(K_EXCL_END_INCL_START if p.kind == K_EXCL_END else K_INCL_END_EXCL_START, p.ck, p.close, u.open)
        else: joined.append(u)
    return joined

def semantic_model(unfiltereds, pdel=None):
    """-> ({position*2 (+1 for 'just after'): deletion ts covering it}, {row position: (ts, row deletion)}) over a small integer domain"""
    cover = {}; rows = {}; open_dt = None; cur = -1
    def fill(upto):
        nonlocal cur
        for x in range(cur + 1, upto): cover[x] = open_dt[0] if open_dt else None
        cur = upto - 1
    for u in unfiltereds:
        pos = struct.unpack(">i", u.ck[0])[0]
        if isinstance(u, Marker):
            before = u.kind in (K_EXCL_END, K_INCL_START, K_EXCL_END_INCL_START)
            fill(2 * pos if before else 2 * pos + 1)
            open_dt = u.open
        else:
            fill(2 * pos); rows[pos] = (u.ts, u.deletion)
    fill(2 * 200)
    return cover, rows

def test_range_tombstone_merge_randomised():
    rng = random.Random(20240)
    b = Builder(S1, (0, 0, 0))
    for it in range(150):
        nsrc = rng.randint(2, 4); srcs = []
        for _ in range(nsrc):
            us = []; pos = 0; open_dt = None
            while pos < 90:
                pos += rng.randint(1, 6); r = rng.random()
                if open_dt is None and r < 0.35:
                    t = rng.randint(100, 200); us.append(Marker(rng.choice((K_INCL_START, K_EXCL_START)), (I32(pos),), None, (t, t))); open_dt = t
                elif open_dt is not None and r < 0.35:
                    if rng.random() < 0.3:
                        t = rng.randint(100, 200)
                        us.append(Marker(rng.choice((K_EXCL_END_INCL_START, K_INCL_END_EXCL_START)), (I32(pos),), (open_dt, open_dt), (t, t))); open_dt = t
                    else:
                        us.append(Marker(rng.choice((K_INCL_END, K_EXCL_END)), (I32(pos),), (open_dt, open_dt), None)); open_dt = None
                else:
                    us.append(Row((I32(pos),), [Cell(0, rng.randint(90, 210), b"v")], ts=rng.randint(90, 210)))
            if open_dt is not None: us.append(Marker(K_INCL_END, (I32(pos + 1),), (open_dt, open_dt), None))
            srcs.append(us)
        pdel = (rng.randint(100, 160),) * 2 if rng.random() < 0.3 else None
        tables = [b.build([Partition(b"p", us, pdel if i == 0 else None)]) for i, us in enumerate(srcs)]
        parts, _ = compact(tables, gc_grace=10**9, now=10**6)
        got = parts[0]
        want_cover = {}
        models = [semantic_model(us) for us in srcs]
        gcover, grows = semantic_model(got.unfiltereds)
        pd = pdel[0] if pdel else None
        for x in range(0, 2 * 100):
            exp = max([m[0].get(x) or -1 for m in models])
            exp = exp if exp > (pd or -1) else None          # only deletions superseding the partition deletion stay open (:160-168)
            if exp == -1: exp = None
            assert gcover.get(x) == exp, (it, x)
        # merged markers must be well formed: strictly alternating open/close, never two identical consecutive deletions
        open_dt = None
        for u in got.unfiltereds:
            if isinstance(u, Marker):
                if u.close is not None: assert open_dt == u.close
                else: assert open_dt is None
                assert u.open != u.close
                open_dt = u.open
        assert open_dt is None
        # rows: a row survives iff something in it is newer than the deletion covering it
        for pos in range(100):
            srows = [m[1][pos] for m in models if pos in m[1]]
            if not srows: assert pos not in grows; continue
            dele = max(gcover.get(2 * pos) or -1, pd or -1)
            if pos in grows: assert grows[pos][0] == NO_TS or grows[pos][0] > dele

def test_dsl_merge_examples():
    b = Builder(S1, (0, 0, 0))
    a = parse_dsl("5<=[140] 10[150] [140]<20 22<[130] [130]<25 30[150]")
    c = parse_dsl("7<[160] 15[180] [160]<30 40[120]")
    parts, r = compact([b.build([Partition(b"p", a)]), b.build([Partition(b"p", c)])], gc_grace=10**9, now=10**6)
    kinds = [(u.kind, struct.unpack(">i", u.ck[0])[0], u.close and u.close[0], u.open and u.open[0]) if isinstance(u, Marker)
             else ("row", struct.unpack(">i", u.ck[0])[0], u.ts) for u in parts[0].unfiltereds]
    # row 10[150] is shadowed by the [160] range; the [140]/[130] ranges never change the merged open deletion, so their
    # markers vanish (RangeTombstoneMarker.Merger emits only on change); the exclusive end at 30 sorts before row 30
    assert kinds == [(K_INCL_START, 5, None, 140), (K_INCL_END_EXCL_START, 7, 140, 160), ("row", 15, 180), (K_EXCL_END, 30, 160, None),
                     ("row", 30, 150), ("row", 40, 120)]
    assert r.stats["total_source_rows"] == 6 + 1      # merged, non-null unfiltereds entering the purger plus the partition's static-row step (Purger.applyToStatic/applyToRow/applyToMarker -> updateProgress)

def test_builder_output_is_identity_under_oracle():
    """independent writer vs oracle writer: compacting a single table with nothing purgeable reproduces its bytes"""
    rng = random.Random(4)
    s = Schema(["LongType", "UTF8Type"], [("a", "LongType"), ("b", "UTF8Type"), ("c", "Int32Type")])
    parts = []
    for k in range(60):
        us = []
        for ck in sorted({(rng.randint(-5, 5), rng.choice([b"", b"x", b"yy", b"zzzzzzzz" * 40])) for _ in range(rng.randint(1, 30))}):
            cells = [Cell(ci, 1000 + rng.randint(0, 9), v) for ci, v in ((0, struct.pack(">q", rng.getrandbits(40))), (1, rng.choice([b"", b"hello", b"w" * 300])), (2, I32(rng.randint(-9, 9)))) if rng.random() < 0.7]
            ts = 1000 + rng.randint(0, 9) if (rng.random() < 0.8 or not cells) else NO_TS      # no empty rows: those are skipped at read time
            us.append(Row((struct.pack(">q", ck[0]), ck[1]), cells, ts=ts))
        parts.append(Partition(struct.pack(">q", rng.getrandbits(63)) + b"k" * rng.randint(0, 3), us, (5, NOW) if rng.random() < 0.1 else None))
    t = Builder(s, (1000, 0, 0), column_index_size=2048).build(parts)
    r = CompactionTask([t], CompactionController(NOW, 10**9), column_index_size=2048).execute(O.OracleEngine())
    o = r.outputs[0]
    assert o.data == t.data and o.index == t.index and o.compression.chunk_offsets == t.compression.chunk_offsets

def test_empty_rows_are_skipped_at_read_time():
    # UnfilteredSerializer.deserialize :433-447 — a row with no liveness, no deletion and no cells never reaches the merge
    b = Builder(S1, (0, 0, 0))
    t = b.build([Partition(b"k", [Row((I32(1),), []), Row((I32(2),), [Cell(0, 5, b"v")])])])
    parts, r = compact([t], gc_grace=10**9)
    assert [u.ck for u in parts[0].unfiltereds] == [(I32(2),)] and r.stats["total_source_rows"] == 1 + 1

# ---- static rows ----------------------------------------------------------------------------------------------------
# No `oa` fixture under the reference's test data carries static columns, so these pin the oracle against the reference's rules restated:
# mergeStaticRows (UnfilteredRowIterators.java:484-505), PurgeFunction.applyToStatic (:101-106), SortedTableWriter.append/addStaticRow
# (:134-146,188-197), SortedTablePartitionWriter.addStaticRow (:117-126), UnfilteredSerializer.serializeStaticRow (:144-149).
S_ST = Schema(["Int32Type"], [("val", "UTF8Type")], static_columns=[("s1", "UTF8Type"), ("s2", "LongType")])
I64 = lambda v: struct.pack(">q", v)

def test_static_rows_merge_newest_cell_wins_per_column():
    b = Builder(S_ST, (0, 0, 0))
    a = b.build([Partition(b"p", [Row((I32(1),), [Cell(0, 10, b"r")], ts=10)], static=Row((), [Cell(0, 100, b"old"), Cell(1, 300, I64(7))]))])
    c = b.build([Partition(b"p", [Row((I32(2),), [Cell(0, 11, b"q")], ts=11)], static=Row((), [Cell(0, 200, b"new")]))])
    parts, r = compact([a, c], schema=S_ST, gc_grace=10**9)
    st = parts[0].static
    assert [(x.col, x.ts, x.value) for x in st.cells] == [(0, 200, b"new"), (1, 300, I64(7))]
    assert [u.ck for u in parts[0].unfiltereds] == [(I32(1),), (I32(2),)]
    assert r.stats["total_source_rows"] == 2 + 1                      # two rows + the static-row step

def test_static_row_identity_with_promoted_index_and_empty_static_rows():
    rng = random.Random(11); parts = []
    for k in range(40):
        rows = [Row((I32(i),), [Cell(0, 1000 + i, b"v" * rng.randint(0, 200))], ts=1000 + i) for i in range(rng.randint(0, 60))]
        st = None
        if rng.random() < 0.6: st = Row((), [Cell(ci, 900 + rng.randint(0, 50), v) for ci, v in ((0, b"s" * rng.randint(0, 40)), (1, I64(k))) if rng.random() < 0.7])
        if st is not None and not st.cells: st = None
        if not rows and st is None: rows = [Row((I32(0),), [Cell(0, 1000, b"x")], ts=1000)]
        parts.append(Partition(b"key%03d" % k, rows, None, st))
    t = Builder(S_ST, (900, 0, 0), column_index_size=1024).build(parts)
    r = CompactionTask([t], CompactionController(NOW, 10**9), column_index_size=1024).execute(O.OracleEngine())
    o = r.outputs[0]
    assert o.data == t.data and o.index == t.index           # headerLength of the promoted index includes the static row

def test_partition_deletion_shadows_static_cells_even_from_a_single_source():
    b = Builder(S_ST, (0, 0, 0))
    # Row.Merger.merge(partitionDeletion) runs for the static row whatever the fan-in, while the clustered rows of a lone source pass through
    t = b.build([Partition(b"p", [Row((I32(1),), [Cell(0, 50, b"kept: TrivialOneToOne")], ts=50)], deletion=(100, NOW), static=Row((), [Cell(0, 50, b"shadowed"), Cell(1, 150, I64(1))]))])
    parts, _ = compact([t], schema=S_ST, gc_grace=10**9)
    assert [(x.col, x.ts) for x in parts[0].static.cells] == [(1, 150)]
    assert len(parts[0].unfiltereds) == 1 and parts[0].deletion == (100, NOW)

def test_static_only_partition_is_written_and_purged_static_only_partition_is_dropped():
    b = Builder(S_ST, (0, 0, 0))
    live = Partition(b"live", [], static=Row((), [Cell(0, 5, b"x")]))
    dead = Partition(b"dead", [], static=Row((), [Cell.tombstone(0, 5, NOW - 10**6)]))
    t = b.build([live, dead])
    parts, r = compact([t], schema=S_ST, gc_grace=1000)
    assert [p.key for p in parts] == [b"live"] and parts[0].static.cells[0].value == b"x"
    assert r.outputs[0].partitions == 1 and r.outputs[0].rows == 0
    parts, _ = compact([t], schema=S_ST, gc_grace=10**9)               # not yet purgeable: both stay
    assert sorted(p.key for p in parts) == [b"dead", b"live"]

def test_static_columns_missing_in_one_input_and_header_union():
    s_a = Schema(["Int32Type"], [("val", "UTF8Type")], static_columns=[("s2", "LongType")])
    s_none = Schema(["Int32Type"], [("val", "UTF8Type")])
    a = Builder(s_a, (0, 0, 0)).build([Partition(b"p", [Row((I32(1),), [Cell(0, 1, b"a")], ts=1)], static=Row((), [Cell(0, 9, I64(42))]))])
    n = Builder(s_none, (0, 0, 0)).build([Partition(b"p", [Row((I32(1),), [Cell(0, 2, b"b")], ts=2)]), Partition(b"q", [Row((I32(1),), [Cell(0, 2, b"c")], ts=2)])])
    c = Builder(S_ST, (0, 0, 0)).build([Partition(b"p", [], static=Row((), [Cell(0, 3, b"s1")]))])
    parts, _ = compact([a, n, c], schema=S_ST, gc_grace=10**9)
    by = {p.key: p for p in parts}
    assert [(x.col, x.value) for x in by[b"p"].static.cells] == [(0, b"s1"), (1, I64(42))] and by[b"q"].static is None
    assert by[b"p"].unfiltereds[0].cells[0].value == b"b"

def test_static_row_counts_in_statistics():
    b = Builder(S_ST, (0, 0, 0))
    t = b.build([Partition(b"p", [Row((I32(1),), [Cell(0, 10, b"r")], ts=10)], static=Row((), [Cell(0, 100, b"x"), Cell(1, 300, I64(7))])),
                 Partition(b"q", [Row((I32(1),), [Cell(0, 10, b"r")], ts=10)])])
    r = CompactionTask([t], CompactionController(NOW, 10**9), with_metadata=True).execute(O.OracleEngine())
    st = r.outputs[0].stats
    # Rows.collectStats runs for the non-empty static row: one more row, two more cells / columns set (SortedTableWriter.addStaticRow :188-197)
    assert st["total_rows"] == 3 and st["total_cells"] == 4 and st["total_columns_set"] == 4 and st["max_timestamp"] == 300
