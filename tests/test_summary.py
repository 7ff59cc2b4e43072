"""Summary.db positions (the seeds of the parallel Index.db walk): the parser against the golden `oa` files and against a blob
laid out per IndexSummary.IndexSummarySerializer.serialize (S/io/sstable/indexsummary/IndexSummary.java:401-423)."""
import os, struct, glob
import numpy as np
from cassandra_b200.io import sstable as S

def test_golden_summaries_point_at_index_entries(golden_dir):
    bases = [d + "oa-1-big-" for d in glob.glob(os.path.join(golden_dir, "oa", "legacy_tables", "*", ""))]
    assert bases
    for base in bases:
        sp = S.parse_summary_positions(open(base + "Summary.db", "rb").read())
        walk = S.index_summary_positions(open(base + "Index.db", "rb").read(), 128)
        assert list(sp) == list(walk) and sp[0] == 0

def test_summary_layout_round_trip():
    keys = [b"k%03d" % i * (1 + i % 3) for i in range(7)]; positions = [0, 131, 4099, 70000, 2**31 + 5, 2**33, 2**40 + 1]
    entries = b"".join(k + struct.pack("<q", p) for k, p in zip(keys, positions))
    offs = np.cumsum([0] + [len(k) + 8 for k in keys[:-1]]) + 4 * len(keys)
    blob = struct.pack(">iiqii", 128, len(keys), 4 * len(keys) + len(entries), 128, len(keys)) + offs.astype("<i4").tobytes() + entries
    blob += struct.pack(">i", 4) + b"frst" + struct.pack(">i", 4) + b"last"          # first/last key trailer, ignored by the parser
    assert list(S.parse_summary_positions(blob)) == positions

def test_builder_and_synth_attach_summaries():
    import synth
    raw = synth.generate_raw(0, 0, 2, 99, 3000)
    want = S.index_summary_positions(raw["index"], 128)
    assert list(raw["summary"]) == list(want) and len(want) > 5
