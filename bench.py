#!/usr/bin/env python
"""bench.py — compaction MB/s (uncompressed in) + rows merged/s on N x B200 (BASELINE.json metric).

A step = one full compaction (b200c_compact: K1 decompress+CRC verify -> K2 index scan -> K3 partition merge -> K4 row merge /
purge / serialise -> K5 LZ4 + CRC32) of one batch of synthetic input SSTables.
  workload (N=1)  BASELINE.json configs[1]: STCS, 16 SSTables x 1 GiB (uncompressed), LZ4, chunk 16 KiB, schema N, seed 0xCA550002.
  N>1             weak scaling: every rank compacts its own 16 x 1 GiB token-range shard (its own seed); no data-path collective,
                  one NCCL broadcast of the run manifest + one all-reduce(max) of the step time.
  value           whole-job MB/s with the inputs (compressed Data.db, Index.db, chunk offsets) already resident in HBM.
  e2e             the same metric through the C ABI with HOST buffers: pinned host -> device copies of every input and the
                  device -> host read-back of Data.db/Index.db/offsets are inside the timed region.
  roofline        the dominant kernel stage (CUDA-event time from the engine's own stream) against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline    the CPU oracle (C++ restatement of the reference algorithm, 1 thread = 1 compaction task as in the reference)
                  on a bounded 1/32-scale sample of the same workload shape, timed on this box's host cores.
--impl reference  the CPU oracle with all host threads (one independent compaction task per thread), same metric/unit.
Input synthesis (synth/) never touches oracle/: inputs are compressed by the engine's own K5 kernels.
"""
import argparse, ctypes as C, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
NOW = 1700000000
METRIC = "compaction MB/s (uncompressed in)"

def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)

def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try: return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
        except Exception: pass
    return 6650.0, "fallback"

class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    def __init__(self, gpu):
        super().__init__(daemon=True); self.gpu = gpu; self.samples = []; self.reasons = set(); self._stop_ev = threading.Event(); self.max_mhz = None
    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:6]):
                    if "Active" in v and "Not" not in v: self.reasons.add(name)
            except Exception:
                pass
            self._stop_ev.wait(0.2)
    def stop(self):
        self._stop_ev.set(); self.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}

# ---------------------------------------------------------------------------------------------------------------------------
def make_inputs_gpu(ctx, schema, nsst, seed, per_sstable_bytes, p, threads, pinned=True, rpp=1000):
    """Synthesises nsst input SSTables and compresses them with the engine's own K5 (b200c_compress_chunks). Returns SSTable
    objects whose data/index/offsets live in pinned host memory (torch tensors kept alive in .hold)."""
    import numpy as np, torch, synth
    from cassandra_b200 import native
    from cassandra_b200.io.sstable import SSTable
    from cassandra_b200.io.compress import CompressionMetadata
    L = native.lib()
    universe = synth.universe_for(schema, per_sstable_bytes, p, rpp)
    tabs = []
    for s in range(nsst):
        t0 = time.time()
        raw = synth.generate_raw(schema, s, nsst, seed, universe, p, rows_per_partition=rpp, threads=threads)
        n = len(raw["stream"]); cap = L.b200c_compress_bound(native.COMP_LZ4, n, 16384); nch = L.b200c_chunk_count(n, 16384)
        out = np.empty(cap, dtype=np.uint8); offs = np.zeros(max(nch, 1), dtype=np.uint64); out_len = C.c_uint64(); dig = C.c_uint32()
        ctx.check(L.b200c_compress_chunks(ctx.handle, native.COMP_LZ4, raw["stream"].ctypes.data, n, 16384, native.INT32_MAX,
                                          out.ctypes.data, cap, C.byref(out_len), offs.ctypes.data, C.byref(dig), 0))
        def pin(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            return t.pin_memory() if pinned else t
        d = pin(out[:out_len.value]); ix = pin(np.frombuffer(raw["index"], dtype=np.uint8)); of = pin(offs[:nch].view(np.int64))
        sm = pin(raw["summary"].view(np.int64))                     # Summary.db sample positions (every 128th Index.db entry)
        meta = CompressionMetadata("LZ4Compressor", 16384, native.INT32_MAX, n, [])
        t = SSTable(None, None, meta, raw["stats"], raw["stats"], synth.SCHEMAS[schema]["clustering"], synth.SCHEMAS[schema]["columns"], generation=s)
        t.hold = (d, ix, of); t.summary = sm; t.nchunks = nch; t.partitions = raw["partitions"]; t.rows = raw["rows"]
        tabs.append(t)
        log("input %d/%d: %.1f MiB uncompressed -> %.1f MiB, %d partitions (%.1fs)" % (s + 1, nsst, n / 2**20, out_len.value / 2**20, raw["partitions"], time.time() - t0))
        del raw, out
    return tabs

def build_manifest(tabs, schema, device_copies=None):
    """b200c_manifest over the inputs; device_copies = list of (data_ptr, index_ptr, offs_ptr) to use instead of the host tensors."""
    import synth
    from cassandra_b200 import native
    from cassandra_b200.io import sstable as sst
    from cassandra_b200.db.compaction import merged_encoding_stats, INT64_MIN, INT64_MAX
    m = native.Manifest(); m.abi_version = native.ABI_VERSION; m.ninputs = len(tabs)
    arr = (native.Input * len(tabs))()
    for k, t in enumerate(tabs):
        d, ix, of = t.hold
        a = arr[k]
        if device_copies: a.data, a.index, a.chunk_offsets, a.summary_positions = device_copies[k]
        else: a.data, a.index, a.chunk_offsets, a.summary_positions = d.data_ptr(), ix.data_ptr(), of.data_ptr(), t.summary.data_ptr()
        a.nsummary = t.summary.numel()
        a.data_len = d.numel(); a.index_len = ix.numel(); a.nchunks = t.nchunks; a.data_length = t.compression.data_length
        a.compressor = native.COMP_LZ4; a.chunk_len = 16384; a.max_compressed_len = native.INT32_MAX
        a.ncolumns = len(t.regular_columns)
        for ci in range(a.ncolumns): a.column_map[ci] = ci
        a.header_stats.min_timestamp, a.header_stats.min_local_deletion_time, a.header_stats.min_ttl = t.header_stats
    m.inputs = arr
    sc = synth.SCHEMAS[schema]
    m.nclustering = len(sc["clustering"])
    for k, ty in enumerate(sc["clustering"]): m.clustering[k].type, m.clustering[k].fixed_len = sst.type_class(ty)
    m.ncolumns = len(sc["columns"])
    for k, (_, ty) in enumerate(sc["columns"]): m.columns[k].type, m.columns[k].fixed_len = sst.type_class(ty)
    m.out_stats.min_timestamp, m.out_stats.min_local_deletion_time, m.out_stats.min_ttl = merged_encoding_stats(tabs)
    m.out_compressor = native.COMP_LZ4; m.out_chunk_len = 16384; m.out_max_compressed_len = native.INT32_MAX; m.column_index_size = 65536
    m.now_in_sec = NOW; m.gc_before = NOW - 864000; m.purge_max_timestamp = INT64_MAX
    m.token_lo, m.token_hi = INT64_MIN, INT64_MAX
    m._keep = arr
    return m

def run_b200(args):
    import numpy as np, torch
    import torch.distributed as dist
    from cassandra_b200 import native
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    L = native.lib()
    ctx = native.Context(local)
    nsst, per = args.sstables, int(args.sstable_mib * 2**20)
    seed = 0xCA550002 + 1000 * rank
    # the run manifest (workload shape) is broadcast once over NCCL so every rank compacts the same shape of shard
    from cassandra_b200 import parallel
    shape = parallel.broadcast_manifest(dict(nsst=nsst, per=per, steps=args.steps, warmup=args.warmup, now=NOW) if rank == 0 else None)
    nsst, per = shape["nsst"], shape["per"]
    threads = max(1, (os.cpu_count() or 8) // max(1, world))
    t0 = time.time()
    schema = 0 if args.schema == "N" else 1
    tabs = make_inputs_gpu(ctx, schema, nsst, seed, per, 0.5, threads, rpp=args.rows_per_partition)
    log("rank %d: inputs ready in %.1fs" % (rank, time.time() - t0))
    u_in = sum(t.compression.data_length for t in tabs); c_in = sum(t.hold[0].numel() for t in tabs); i_in = sum(t.hold[1].numel() for t in tabs)

    # host (e2e) manifest + outputs in pinned memory
    m_host = build_manifest(tabs, schema)
    cap_d = L.b200c_compress_bound(native.COMP_LZ4, u_in, 16384); cap_i = i_in + (1 << 20); cap_c = u_in // 16384 + 16
    ho = (torch.empty(cap_d, dtype=torch.uint8).pin_memory(), torch.empty(cap_i, dtype=torch.uint8).pin_memory(), torch.empty(cap_c, dtype=torch.int64).pin_memory())
    def result_for(bufs):
        res = native.Result(); outs = (native.Output * 1)(); o = outs[0]
        o.data, o.data_cap, o.index, o.index_cap, o.chunk_offsets, o.chunk_cap = bufs[0].data_ptr(), cap_d, bufs[1].data_ptr(), cap_i, bufs[2].data_ptr(), cap_c
        res.noutputs_cap = 1; res.outputs = outs; res._keep = outs
        return res
    # device-resident copies of the inputs and outputs (value)
    dev_in = [(t.hold[0].cuda(), t.hold[1].cuda(), t.hold[2].cuda(), t.summary.cuda()) for t in tabs]
    m_dev = build_manifest(tabs, schema, [(a.data_ptr(), b.data_ptr(), c.data_ptr(), d_.data_ptr()) for a, b, c, d_ in dev_in])
    do = (torch.empty(cap_d, dtype=torch.uint8, device="cuda"), torch.empty(cap_i, dtype=torch.uint8, device="cuda"), torch.empty(cap_c, dtype=torch.int64, device="cuda"))

    def step(dev):
        res = result_for(do if dev else ho)
        ctx.check(L.b200c_compact(ctx.handle, C.byref(m_dev if dev else m_host), C.byref(res), 1 if dev else 0), res.corruption)
        return res

    def timed(dev, steps, sampler=None):
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        if sampler: sampler.start()
        launches0 = ctx.total_kernel_launches
        t = time.perf_counter(); kms = 0.0; stages = [0.0] * 6; last = None
        for _ in range(steps):
            last = step(dev); kms += last.kernel_ms
            for i, v in enumerate(ctx.last_stage_ms()): stages[i] += v
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        clocks = sampler.stop() if sampler else None
        return parallel.max_over_ranks(dt, device="cuda"), kms / steps, [s / steps for s in stages], last, ctx.total_kernel_launches - launches0, clocks

    for _ in range(args.warmup): step(True)
    dt, kms, stages, last, launches, clocks = timed(True, args.steps, ClockSampler(local))
    u_out, c_out, i_out = int(last.bytes_written), int(last.outputs[0].data_len), int(last.outputs[0].index_len)
    rows = int(last.total_source_rows); parts_in = int(last.input_partitions)
    total_in_all = u_in * world            # every rank has the same shape (weak scaling); rank-local sizes differ by < 0.1 %
    value = total_in_all * args.steps / dt / 1e6
    # e2e (host buffers) — fewer steps are fine, the copies dominate
    step(False)
    e_steps = max(1, min(args.steps, 3))
    edt, ekms, estages, elast, _, _ = timed(False, e_steps)
    e2e = total_in_all * e_steps / edt / 1e6
    sweep = {}
    for rr in [x for x in args.e2e_ranges.split(",") if x]:
        os.environ["B200C_RANGES"] = rr
        step(False)
        sdt, _, sst, _, _, _ = timed(False, e_steps)
        sweep[rr] = {"ms_per_step": round(sdt / e_steps * 1e3, 2), "stage_ms": [round(x, 1) for x in sst]}
        del os.environ["B200C_RANGES"]
    ab = None
    if args.ab_env:
        k_, v_ = args.ab_env.split("=", 1); os.environ[k_] = v_
        step(True); step(False)
        adt, _, ast_, _, _, _ = timed(True, args.steps)
        bdt, _, bst_, _, _, _ = timed(False, e_steps)
        ab = {"env": args.ab_env, "value": round(total_in_all * args.steps / adt / 1e6, 1), "e2e": round(total_in_all * e_steps / bdt / 1e6, 1),
              "stage_ms": [round(x, 1) for x in ast_], "e2e_stage_ms": [round(x, 1) for x in bst_]}
        del os.environ[k_]
    h2d = c_in + i_in + 8 * sum(t.nchunks for t in tabs) + 8 * sum(t.summary.numel() for t in tabs); d2h = c_out + i_out + 8 * int(elast.outputs[0].nchunks)

    # roofline of the dominant stage, algorithmic bytes per SURVEY §8(d): every compressed byte read once, every uncompressed byte
    # produced once, merged stream written once and compressed once
    names = ["K1 decompress+verify", "K2 index scan", "K3 partition merge", "K4 merge+purge+serialise", "K4 gather+index", "K5 compress+crc+pack"]
    alg = [c_in + u_in, i_in + 26 * parts_in, 26 * parts_in + 20 * parts_in, u_in + u_out, 2 * u_out + i_out, u_out + c_out]
    dom = max(range(6), key=lambda i: stages[i])
    peak, which = peaks()
    b_alg = c_in + i_in + u_in + u_out + c_out + i_out
    roof = {"bound": "hbm", "kernel": names[dom], "achieved": round(alg[dom] / (stages[dom] / 1e3) / 1e9, 2), "peak": peak, "unit": "GB/s",
            "frac": round(alg[dom] / (stages[dom] / 1e3) / 1e9 / peak, 5), "traffic": None, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)" if which == "measured" else which,
            "stage_ms": {n: round(s, 3) for n, s in zip(names, stages)}, "kernel_ms_per_step": round(kms, 3),
            "pipeline_achieved_gbs": round(b_alg / (kms / 1e3) / 1e9, 2), "pipeline_frac": round(b_alg / (kms / 1e3) / 1e9 / peak, 5),
            "algorithmic_bytes_per_step": b_alg}
    line = {"metric": METRIC, "value": round(value, 1), "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "STCS %d SSTables x %d MiB (uncompressed), LZ4 16 KiB chunks, schema %s, 1 output; %s" %
                       (nsst, per >> 20, args.schema if args.schema == "N" else "W (%d rows/partition)" % args.rows_per_partition,
                        "BASELINE.json configs[1]" if (nsst, per >> 20, args.schema) == (16, 1024, "N") else "NOT configs[1] (16 x 1024 MiB, schema N)"),
                       "per_gpu_uncompressed_in_bytes": u_in, "l2": "inputs (%.1f GB/step) larger than L2" % ((c_in + u_in) / 1e9),
                       "parallelism": "token-range shard per GPU, no data-path collective", "now_in_sec": NOW, "gc_grace": 864000, "seed": seed},
            "rows_merged_per_s": round(rows * world * args.steps / dt, 0), "input_partitions_per_step": parts_in,
            "merged_row_counts": [int(x) for x in last.merged_row_counts[:nsst]],
            "bytes": {"u_in": u_in, "c_in": c_in, "index_in": i_in, "u_out": u_out, "c_out": c_out, "index_out": i_out},
            "e2e": {"value": round(e2e, 1), "unit": "MB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": round(edt / e_steps * 1e3, 2),
                    "device_ms_per_step": round(ekms, 2), "stage_ms": {n: round(s_, 2) for n, s_ in zip(names, estages)}},
            "e2e_ranges_sweep": sweep, "ab": ab, "gpu_launches": int(launches), "index_slow_path_inputs": int(last.index_slow_path_inputs), "clocks": clocks, "roofline": roof}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_sample(1, args.cpu_sample_mib, 1)
    if rank == 0:
        emit(line)
    ctx.close()
    if world > 1: dist.destroy_process_group()

# ---------------------------------------------------------------------------------------------------------------------------
def cpu_sample(threads, sstable_mib, steps, warmup=0):
    """The CPU oracle (one single-threaded compaction task per thread, like the reference's CompactionExecutor) on a bounded sample
    with the workload's shape: 16 input SSTables per task, schema N, LZ4. Inputs are compressed by the oracle's own CPU LZ4."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O, synth
    from synth_util import synth_tables
    from cassandra_b200.db.compaction import CompactionTask, CompactionController
    universe = synth.universe_for(0, int(sstable_mib * 2**20), 0.5)
    from concurrent.futures import ThreadPoolExecutor
    def make_task(t):            # inputs are synthesised and LZ4-compressed (oracle codec) in parallel; ctypes releases the GIL
        tabs = []
        for s_ in range(16):
            raw = synth.generate_raw(0, s_, 16, 0xCA550002 + 7777 * (t + 1), universe, 0.5, threads=1)
            from synth_util import oracle_compress
            tb = synth.make_sstable(raw, 0, lambda st, cl: oracle_compress(st, cl), "LZ4Compressor", generation=s_)
            tabs.append(tb)
        return CompactionTask(tabs, CompactionController(NOW))
    with ThreadPoolExecutor(max_workers=min(threads, os.cpu_count() or 8)) as ex:
        tasks = list(ex.map(make_task, range(threads)))
    total = sum(i.compression.data_length for task in tasks for i in task.inputs)
    rows = [0] * threads
    def work(k):
        r = tasks[k].execute(O.OracleEngine()); rows[k] = r.stats["total_source_rows"]
    def one_step():
        th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        t0 = time.perf_counter()
        for x in th: x.start()
        for x in th: x.join()
        return time.perf_counter() - t0
    for _ in range(warmup): one_step()
    dt = sum(one_step() for _ in range(steps))
    return {"value": round(total * steps / dt / 1e6, 1), "unit": "MB/s", "cores": threads, "kind": "port",
            "sample": "%d task(s) x 16 SSTables x %g MiB (1/%d scale of configs[1]), C++ oracle, 1 thread per task" % (threads, sstable_mib, int(1024 / sstable_mib)),
            "rows_merged_per_s": round(sum(rows) * steps / dt, 0), "ms_per_step": round(dt / steps * 1e3, 1)}

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0: return
    threads = args.ref_threads or (os.cpu_count() or 8)          # every host thread: one independent compaction task each
    cb = cpu_sample(threads, args.ref_sample_mib, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "STCS 16 SSTables per task, LZ4 16 KiB chunks, schema N; CPU reference arm = C++ restatement of the reference "
                       "algorithm (the JVM cannot run in this image), one compaction task per host thread"},
            "rows_merged_per_s": cb["rows_merged_per_s"], "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)

_REAL_STDOUT = None
def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n"); out.flush()

def main():
    # the contract is ONE JSON line on stdout: libraries (NCCL prints its version banner there) are diverted to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sstables", type=int, default=16)
    ap.add_argument("--sstable-mib", type=float, default=1024.0, help="uncompressed size of each input (configs[1]: 1024)")
    ap.add_argument("--e2e-ranges", default="", help="development aid: also time the e2e path with B200C_RANGES forced to each of these comma-separated values")
    ap.add_argument("--ab-env", default="", help="development aid: NAME=VALUE; after the normal measurement, time both legs again with this environment variable set")
    ap.add_argument("--schema", default="N", choices=["N", "W"], help="N: narrow rows (configs[1..3]); W: wide time-series partitions (configs[4] shape)")
    ap.add_argument("--rows-per-partition", type=int, default=1000)
    ap.add_argument("--cpu-sample-mib", type=float, default=32.0)
    ap.add_argument("--ref-sample-mib", type=float, default=16.0)
    ap.add_argument("--ref-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200": log("warning: fewer than 3 warm-up steps")
    if args.impl == "reference": run_reference(args)
    else: run_b200(args)

if __name__ == "__main__":
    main()
