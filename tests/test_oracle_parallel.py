"""The multi-threaded CPU oracle (oracle/parallel.cc: one compaction cut into token ranges, one oracle thread per range, stitched and
compressed chunk-parallel) must reproduce the single-threaded oracle byte for byte — it is what bench.py times as the all-cores CPU
baseline and what verifies the GPU output of the full-size benchmark."""
import ctypes as C, numpy as np, pytest
import oracle_lib as O
from synth_util import synth_tables
from cassandra_b200 import native
from cassandra_b200.db.compaction import CompactionTask, CompactionController

NOW = 1700000000

def run_parallel(task, threads, ranges, max_ranges=0):
    m = task.build_manifest()
    total = sum(i.compression.data_length for i in task.inputs)
    d = np.empty(total * 2 + (1 << 20), dtype=np.uint8); ix = np.empty(sum(len(i.index) for i in task.inputs) * 2 + (1 << 16), dtype=np.uint8)
    co = np.zeros(total // task.compression.chunk_length + 16, dtype=np.uint64)
    res = native.Result(); outs = (native.Output * 1)(); o = outs[0]
    o.data, o.data_cap, o.index, o.index_cap, o.chunk_offsets, o.chunk_cap = d.ctypes.data, len(d), ix.ctypes.data, len(ix), co.ctypes.data, len(co)
    res.noutputs_cap = 1; res.outputs = outs
    L = O.lib(); L.orc_compact_parallel.restype = C.c_int
    L.orc_compact_parallel.argtypes = [C.POINTER(native.Manifest), C.POINTER(native.Result), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_char_p, C.c_int]
    err = C.create_string_buffer(256); tm = (C.c_double * 6)(); hi = C.c_int64()
    rc = L.orc_compact_parallel(C.byref(m), C.byref(res), threads, ranges, max_ranges, tm, C.byref(hi), err, 256)
    assert rc == 0, err.value
    return (bytes(d[:o.data_len]), bytes(ix[:o.index_len]), [int(x) for x in co[:o.nchunks]], int(o.digest), int(o.partitions), int(o.rows),
            {k: (int(getattr(res, k)) if k != "merged_row_counts" else [int(x) for x in res.merged_row_counts[:len(task.inputs)]]) for k in
             ("bytes_read", "bytes_in_range", "bytes_written", "total_source_rows", "input_partitions", "merged_row_counts")}, hi.value)

@pytest.mark.parametrize("schema,n,universe,rpp,threads,ranges", [(0, 6, 30000, 0, 4, 7), (0, 16, 6000, 0, 8, 32), (1, 3, 60, 500, 3, 5), (0, 2, 2000, 0, 1, 1)])
def test_parallel_oracle_equals_single_threaded_oracle(schema, n, universe, rpp, threads, ranges):
    tabs = synth_tables(schema, n, 0x9A7 + n, universe, rows_per_partition=rpp)
    for g, t in enumerate(tabs): t.generation = g
    task = CompactionTask(tabs, CompactionController(NOW))
    want = task.execute(O.OracleEngine()); w = want.outputs[0]
    data, index, offs, digest, parts, rows, stats, hi = run_parallel(CompactionTask(tabs, CompactionController(NOW)), threads, ranges)
    assert data == w.data and index == w.index and offs == w.compression.chunk_offsets and digest == w.digest
    assert (parts, rows) == (w.partitions, w.rows) and hi == (1 << 63) - 1
    for k, v in stats.items(): assert want.stats[k] == v, k

def test_parallel_oracle_sample_is_a_token_range_prefix():
    tabs = synth_tables(0, 4, 0x9A8, 20000)
    for g, t in enumerate(tabs): t.generation = g
    data, index, offs, digest, parts, rows, stats, hi = run_parallel(CompactionTask(tabs, CompactionController(NOW)), 4, 16, max_ranges=5)
    assert hi < (1 << 63) - 1
    want = CompactionTask(tabs, CompactionController(NOW), token_range=(-(1 << 63), hi)).execute(O.OracleEngine())
    w = want.outputs[0]
    assert data == w.data and index == w.index and digest == w.digest and stats["bytes_in_range"] == want.stats["bytes_in_range"] < want.stats["bytes_read"]
