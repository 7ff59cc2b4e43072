"""smoke(): one small 4-way compaction on cuda:0 checked byte for byte against the CPU oracle."""
def smoke_compaction(ctx):
    import oracle_lib as O
    from synth_util import synth_tables
    from cassandra_b200.db.compaction import CompactionTask, CompactionController, GpuEngine
    tabs = synth_tables(0, 4, 0xCA550001, 6000)
    ctl = CompactionController(1700000000)
    want = CompactionTask(tabs, ctl).execute(O.OracleEngine()).outputs[0]
    got = CompactionTask(tabs, ctl).execute(GpuEngine(ctx)).outputs[0]
    assert got.data == want.data and got.index == want.index and got.digest == want.digest, "GPU compaction differs from the oracle"
    assert got.compression.chunk_offsets == want.compression.chunk_offsets
