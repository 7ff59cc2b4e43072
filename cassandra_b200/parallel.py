"""Multi-GPU plumbing of the compaction path: the path shards by token range with NO data-path collective.
One process per GPU (torch.distributed, NCCL on the GPU box, gloo in CPU tests). Collectives used, all tiny:
  * one broadcast of the run manifest (shape / splitters / determinism inputs) from rank 0,
  * one all-gather of per-shard counters (partitions, bytes) for reporting,
  * one all-reduce(MAX) of the step time (device-timed, max over ranks).
Reference precedent for ranged compaction: AbstractCompactionStrategy.getScanners(sstables, ranges)
(S/db/compaction/AbstractCompactionStrategy.java:247-269) and UCS ShardedCompactionWriter (S/db/compaction/unified/ShardedCompactionWriter.java:65-82)."""
INT64_MIN, INT64_MAX = -(1 << 63), (1 << 63) - 1

def shard_token_ranges(world_size: int):
    """Equal-width (lo, hi] token ranges covering the Murmur3 ring; shard 0 starts at MIN (inclusive by the engine's convention)."""
    span = 1 << 64
    cuts = [INT64_MIN + (span * k) // world_size for k in range(world_size)] + [INT64_MAX]
    return [(cuts[k] if k else INT64_MIN, cuts[k + 1] if k + 1 < world_size else INT64_MAX) for k in range(world_size)]

def weighted_token_ranges(sample_tokens, world_size: int):
    """Splitters that balance the number of sampled partitions per shard (host picks G-1 splitters from Summary/Index samples)."""
    t = sorted(sample_tokens)
    if not t: return shard_token_ranges(world_size)
    cuts = [INT64_MIN] + [t[(len(t) * k) // world_size] for k in range(1, world_size)] + [INT64_MAX]
    return [(cuts[k], cuts[k + 1]) for k in range(world_size)]

def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None

def broadcast_manifest(obj, src=0):
    """rank `src` decides the run manifest (a small picklable dict); everyone gets the same copy."""
    d = _dist()
    if d is None: return obj
    box = [obj if d.get_rank() == src else None]
    d.broadcast_object_list(box, src=src)
    return box[0]

def all_gather_counters(counters: dict):
    d = _dist()
    if d is None: return [counters]
    out = [None] * d.get_world_size()
    d.all_gather_object(out, counters)
    return out

def max_over_ranks(seconds: float, device=None) -> float:
    d = _dist()
    if d is None: return seconds
    import torch
    t = torch.tensor([seconds], dtype=torch.float64, device=device or "cpu")
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t[0])
