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
