"""Pins the CPU oracle's compaction path against the reference's golden `oa` SSTables: identity compaction of a table written
by a real Cassandra release must reproduce Data.db, Index.db, CompressionInfo.db and Digest.crc32 byte for byte (SURVEY §8c:
the primary parity gate; fixtures were written with column_index_size 4 KiB from test/conf/cassandra.yaml)."""
import os, pytest
import oracle_lib as O
from cassandra_b200.io.sstable import SSTable
from cassandra_b200.db.compaction import CompactionTask, CompactionController

def _golden(golden_dir, name):
    return os.path.join(golden_dir, "oa", "legacy_tables", name, "oa-1-big-")

@pytest.mark.parametrize("name", ["legacy_oa_simple", "legacy_oa_clust"])
def test_identity_compaction_reproduces_golden_files(golden_dir, name):
    base = _golden(golden_dir, name)
    s = SSTable.open(base)
    task = CompactionTask([s], CompactionController(now_in_sec=1700000000), column_index_size=4096)
    r = task.execute(O.OracleEngine())
    assert len(r.outputs) == 1
    comp = r.outputs[0].components()
    for c in ("Data.db", "Index.db", "CompressionInfo.db", "Digest.crc32"):
        assert comp[c] == open(base + c, "rb").read(), c
    assert r.stats["bytes_read"] == s.compression.data_length == r.stats["bytes_written"]

def test_self_merge_is_idempotent(golden_dir):
    """merging a table with itself (two identical inputs) yields the same bytes: reconcile picks equal cells, markers collapse"""
    base = _golden(golden_dir, "legacy_oa_clust")
    a, b = SSTable.open(base, 1), SSTable.open(base, 2)
    r = CompactionTask([a, b], CompactionController(1700000000), column_index_size=4096).execute(O.OracleEngine())
    assert r.outputs[0].components()["Data.db"] == open(base + "Data.db", "rb").read()
    assert r.outputs[0].components()["Index.db"] == open(base + "Index.db", "rb").read()
    assert r.stats["merged_row_counts"] == [0, 5]

# ---- round 2: partitioner order, DateType, lazy chunk flush at the LCS switch, per-token-range purge table -------------------------
import struct, random
from sstable_builder import Schema, Builder, Partition, Row, Cell, decode_stream
from cassandra_b200 import native
NOW = 1700000000

def test_wrong_partition_order_is_rejected_whatever_the_file_size(golden_dir):
    """the golden `oa` tables are ByteOrderedPartitioner files (their Statistics.db says so) and compact as such; declared as
    Murmur3 files their 5 keys are out of token order and must be refused, not merged (no small-file exemption)"""
    s = SSTable.open(_golden(golden_dir, "legacy_oa_simple"))
    assert s.partitioner.endswith("ByteOrderedPartitioner")
    s.partitioner = "org.apache.cassandra.dht.Murmur3Partitioner"
    with pytest.raises(native.CorruptSSTableError):
        CompactionTask([s], CompactionController(NOW), column_index_size=4096).execute(O.OracleEngine())
    s.partitioner = "org.apache.cassandra.dht.RandomPartitioner"
    with pytest.raises(native.UnsupportedError):
        CompactionTask([s], CompactionController(NOW)).build_manifest()

def datetype_tables():
    """DateType compares as unsigned bytes (ComparisonType.BYTE_ORDER): a pre-1970 value (sign bit set) sorts AFTER post-1970 ones"""
    sc = Schema(["DateType"], [("v", "UTF8Type")]); b = Builder(sc, (0, 0, 0))
    d = lambda ms: struct.pack(">q", ms)
    t1 = b.build([Partition(b"k", [Row((d(5),), [Cell(0, 10, b"a")], ts=10), Row((d(-7),), [Cell(0, 10, b"old")], ts=10)])])
    t2 = b.build([Partition(b"k", [Row((d(1),), [Cell(0, 11, b"b")], ts=11), Row((d(-7),), [Cell(0, 12, b"new")], ts=12), Row((d(-2),), [Cell(0, 11, b"c")], ts=11)])])
    return sc, [t1, t2], [d(1), d(5), d(-7), d(-2)]

def test_datetype_clustering_is_unsigned_byte_order():
    sc, tabs, want = datetype_tables()
    assert native.TYPE_FIXED_BYTES == __import__("cassandra_b200.io.sstable", fromlist=["type_class"]).type_class("DateType")[0]
    r = CompactionTask(tabs, CompactionController(0, 0)).execute(O.OracleEngine())
    from synth_util import decompress_output
    parts = decode_stream(sc, decompress_output(r.outputs[0]), (0, 0, 0))
    assert [u.ck[0] for u in parts[0].unfiltereds] == want
    assert [u.cells[0].value for u in parts[0].unfiltereds] == [b"b", b"a", b"new", b"c"]

def lcs_boundary_tables(chunk=256):
    """partitions of exactly `chunk` bytes: every partition ends on a chunk boundary, the case where eager and lazy flushing differ"""
    sc = Schema([], [("v", "BytesType")]); b = Builder(sc, (0, 0, 0))
    rng = random.Random(5); parts = []
    for k in range(40):
        key = b"key%05d" % k
        # partition = 2 + len(key) + 1 (LIVE) + row + 1 (end). row = flags(1) + vint(size)(1..2) + vint(prev)(1) + ts vint(1) + cell flags(1) + vint(len)(1..2) + value
        fixed = 2 + len(key) + 1 + 1
        val_len = chunk - fixed - (1 + 2 + 1 + 1 + 1 + 2)            # sizes >= 128 take a 2-byte vint
        val = bytes(rng.getrandbits(8) for _ in range(val_len))
        parts.append(Partition(key, [Row((), [Cell(0, 5, val)], ts=5)]))
    t = b.build(parts, chunk_length=chunk)
    sizes = []
    p = 0; data = t.uncompressed
    for q in decode_stream(sc, data, (0, 0, 0)): pass
    return sc, t

def test_lcs_switch_counts_only_flushed_chunks():
    """MaxSSTableSizeWriter switches before a partition when getEstimatedOnDiskBytesWritten() (= bytes of chunks already FLUSHED) exceeds
    the limit; a chunk that is exactly full is flushed lazily, by the next write (BufferedDataOutputStreamPlus.write :87-139)"""
    chunk = 256
    sc, t = lcs_boundary_tables(chunk)
    assert len(t.uncompressed) % chunk == 0
    starts = []; p = 0; d = t.uncompressed
    while p < len(d):
        starts.append(p); p += chunk
    assert len(starts) == 40
    limit = 1000
    r = CompactionTask([t], CompactionController(0, 0), max_sstable_bytes=limit).execute(O.OracleEngine())
    # model: the values are random bytes, so every compressed chunk is a stored-literal LZ4 block of known size
    csize = len(O.chunk_compress(O.COMP_LZ4, d[:chunk])) + 4
    want = []; on_disk = 0; in_file = 0; buffered = 0
    for _ in range(40):
        if on_disk > limit: want.append(in_file); in_file = 0; on_disk = 0; buffered = 0       # switch happens before the partition
        # writing `chunk` bytes into a buffer holding `buffered`: flush lazily when full and more bytes arrive
        if buffered == chunk: on_disk += csize; buffered = 0
        buffered += chunk; in_file += 1
    want.append(in_file)
    assert [o.partitions for o in r.outputs] == want
    assert want[0] == 1000 // csize + 2              # eager flushing would have cut one partition earlier

def purge_table_inputs():
    from synth_util import synth_tables
    tabs = synth_tables(0, 4, 0xCA551234, 3000)
    toks = sorted(O.token(struct.pack(">q", 0)) for _ in range(1))
    return tabs

def test_purge_table_equals_per_range_compactions():
    """a purge table {token bound -> threshold} must give the same bytes as compacting each token range with its own threshold"""
    tabs = purge_table_inputs()
    from synth_util import decompress_output
    cuts = [-(1 << 62), 0, 1 << 62]
    thr = [1600000000000000 + 500000000, 1600000000000000 + 2500000000, (1 << 63) - 1]
    whole = CompactionTask(tabs, CompactionController(NOW, overlapping_min_timestamp=1600000000000000 + 1500000000, purge_ranges=list(zip(cuts, thr)))).execute(O.OracleEngine())
    pieces = []; lo = -(1 << 63)
    for hi, t in list(zip(cuts, thr)) + [((1 << 63) - 1, 1600000000000000 + 1500000000)]:
        r = CompactionTask(tabs, CompactionController(NOW, overlapping_min_timestamp=t), token_range=(lo, hi)).execute(O.OracleEngine())
        pieces.append(decompress_output(r.outputs[0])); lo = hi
    assert decompress_output(whole.outputs[0]) == b"".join(pieces)
    flat = CompactionTask(tabs, CompactionController(NOW, overlapping_min_timestamp=1600000000000000 + 1500000000)).execute(O.OracleEngine())
    assert decompress_output(flat.outputs[0]) != decompress_output(whole.outputs[0])

# ---- Filter.db, Summary.db and the Statistics.db reductions, pinned by the golden components ----------------------------------
import struct as _st
from cassandra_b200.io.sstable import parse_statistics, bloom_geometry

def golden_meta_task(golden_dir, name):
    base = _golden(golden_dir, name)
    s = SSTable.open(base)
    hc, words = _st.unpack_from(">ii", open(base + "Filter.db", "rb").read(), 0)
    return base, s, CompactionTask([s], CompactionController(NOW), column_index_size=4096, bloom=(hc, words))

@pytest.mark.parametrize("name", ["legacy_oa_simple", "legacy_oa_clust"])
def test_identity_compaction_reproduces_filter_summary_and_statistics(golden_dir, name):
    """identity compaction of a table written by a real Cassandra release: Filter.db (BloomFilter.add over hash3_x64_128) and Summary.db
    (IndexSummaryBuilder) byte for byte; the MetadataCollector reductions equal to what the golden Statistics.db holds"""
    base, s, task = golden_meta_task(golden_dir, name)
    o = task.execute(O.OracleEngine()).outputs[0]
    assert o.filter == open(base + "Filter.db", "rb").read()
    assert o.summary == open(base + "Summary.db", "rb").read()
    st = parse_statistics(open(base + "Statistics.db", "rb").read())
    for k in ("min_timestamp", "max_timestamp", "min_local_deletion_time", "max_local_deletion_time", "min_ttl", "max_ttl"):
        assert o.stats[k] == st[k], k
    assert o.stats["partition_size_hist"] == st["partition_size_hist"] and o.stats["cells_per_partition_hist"] == st["cells_per_partition_hist"]
    assert o.stats["tombstone_drop_times"] == st["tombstone_drop_times"] == []
    assert (o.first_key, o.last_key) == (b"0", b"4")
    assert o.stats["total_rows"] == o.rows and o.stats["has_partition_level_deletions"] == 0

def test_bloom_geometry_is_filterfactorys():
    assert bloom_geometry(5, 0.01) == (5, 2)                       # the golden Filter.db: 5 keys at fp 0.01
    assert bloom_geometry(1000000, 0.01) == (5, (1000000 * 10 + 20 - 1) // 64 + 1)
    assert bloom_geometry(1000, 0.1) == (3, (1000 * 5 + 20 - 1) // 64 + 1)
    assert bloom_geometry(10, 1.0) == (0, 0)

def test_metadata_of_a_synthetic_compaction_is_consistent():
    from synth_util import synth_tables, decompress_output
    tabs = synth_tables(0, 4, 0x57A7, 3000)
    r = CompactionTask(tabs, CompactionController(NOW), with_metadata=True, min_index_interval=16).execute(O.OracleEngine())
    o = r.outputs[0]; st = o.stats
    assert sum(st["partition_size_hist"]) == o.partitions == sum(st["cells_per_partition_hist"])
    assert st["total_cells"] == st["total_columns_set"] > 0 and st["total_rows"] == o.rows - sum(1 for _ in []) >= st["total_rows"] > 0 and st["total_tombstones"] > 0
    assert st["min_timestamp"] <= st["max_timestamp"] and st["min_ttl"] == 0 < st["max_ttl"]
    assert sum(c for _, c in st["tombstone_drop_times"]) > 0 and all(p % 60 == 0 for p, _ in st["tombstone_drop_times"])
    # Summary.db: every 16th Index.db entry; the filter holds every key
    from test_golden_murmur3 import index_keys, bloom_bytes
    keys = index_keys(o.index)
    assert (o.first_key, o.last_key) == (keys[0], keys[-1])
    n = _st.unpack_from(">i", o.summary, 4)[0]; assert n == (len(keys) + 15) // 16
    hc, words = _st.unpack_from(">ii", o.filter, 0)
    assert o.filter[8:] == bloom_bytes(keys, hc, words)
    regs = st["hll_registers"]; assert len(regs) == 8192 and sum(1 for x in regs if x) > 0.2 * min(len(keys), 8192)
