"""N > 1 host logic on CPU: two and four gloo ranks shard one compaction by token range (oracle engine stands in for the GPU here — this
test is about the plumbing: manifest broadcast, range splitting, counter gather, max-reduce), and the shards must add up to the
unsharded result exactly (partition records are position independent)."""
import os, sys, tempfile, pickle, pytest
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from synth_util import synth_tables, decompress_output
    from cassandra_b200 import parallel
    from cassandra_b200.db.compaction import CompactionTask, CompactionController
    manifest = parallel.broadcast_manifest(dict(seed=0xCA550004, nsst=4, universe=6000, now=1700000000) if rank == 0 else None)
    tabs = synth_tables(0, manifest["nsst"], manifest["seed"], manifest["universe"])
    # splitters as bench.py --gpus N picks them: rank 0 hashes evenly spaced Summary.db sample keys (b200c_token, host function of the
    # product library) and balances the sampled partitions; one broadcast
    cuts = None
    if rank == 0:
        from cassandra_b200 import native
        L = native.lib(); toks = []
        for t in tabs:
            ix = t.index
            for off in [int(x) for x in t.summary_positions[::(7 if world == 2 else 1)]]:      # (a few dozen samples per input: enough for four shards)
                kl = (ix[off] << 8) | ix[off + 1]; key = bytes(ix[off + 2:off + 2 + kl])
                toks.append(L.b200c_token(native.PARTITIONER_MURMUR3, key, kl))
        cuts = parallel.weighted_token_ranges(toks, world)
    cuts = parallel.broadcast_manifest(cuts)
    lo, hi = cuts[rank]
    r = CompactionTask(tabs, CompactionController(manifest["now"]), token_range=(lo, hi)).execute(O.OracleEngine())
    counters = dict(rank=rank, in_range=r.stats["bytes_in_range"], partitions=r.outputs[0].partitions, rows=r.outputs[0].rows, bytes_written=r.stats["bytes_written"])
    gathered = parallel.all_gather_counters(counters)
    tmax = parallel.max_over_ranks(1.0 + rank)
    pickle.dump(dict(stream=decompress_output(r.outputs[0]), gathered=gathered, tmax=tmax, manifest=manifest), open(os.path.join(outdir, "r%d.pkl" % rank), "wb"))
    dist.destroy_process_group()

@pytest.mark.parametrize("world", [2, 4])
def test_token_range_sharding_over_gloo_ranks(world):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, 29533 + world, d), nprocs=world, join=True)
        res = [pickle.load(open(os.path.join(d, "r%d.pkl" % r), "rb")) for r in range(world)]
    assert all(x["manifest"] == res[0]["manifest"] for x in res) and res[0]["manifest"]["seed"] == 0xCA550004
    assert all(x["tmax"] == float(world) for x in res)
    assert [g["rank"] for g in res[-1]["gathered"]] == list(range(world))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from synth_util import synth_tables, decompress_output
    from cassandra_b200.db.compaction import CompactionTask, CompactionController
    tabs = synth_tables(0, 4, 0xCA550004, 6000)
    full = CompactionTask(tabs, CompactionController(1700000000)).execute(O.OracleEngine())
    assert b"".join(x["stream"] for x in res) == decompress_output(full.outputs[0])
    assert sum(g["partitions"] for g in res[0]["gathered"]) == full.outputs[0].partitions
    assert sum(g["rows"] for g in res[0]["gathered"]) == full.outputs[0].rows
    shares = [g["in_range"] for g in res[0]["gathered"]]
    assert sum(shares) == full.stats["bytes_read"] and min(shares) > 0.7 / world * sum(shares)        # the sample-weighted splitter balances the shards

def test_range_helpers():
    from cassandra_b200 import parallel
    for w in (1, 2, 4, 8):
        r = parallel.shard_token_ranges(w)
        assert r[0][0] == parallel.INT64_MIN and r[-1][1] == parallel.INT64_MAX
        assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo < hi for lo, hi in r)
    w = parallel.weighted_token_ranges(list(range(-1000, 1000, 7)), 4)
    assert len(w) == 4 and w[0][0] == parallel.INT64_MIN and w[-1][1] == parallel.INT64_MAX and all(a[1] == b[0] for a, b in zip(w, w[1:]))
