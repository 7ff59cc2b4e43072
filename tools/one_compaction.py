#!/usr/bin/env python
"""Development driver for profilers: builds one workload's inputs (bench.py helpers) and runs N compactions, printing the stage clock.
  python tools/one_compaction.py --workload cfg1 --mib 64 --repeat 2 [--host]
Used under ncu (launch list / --set full of one kernel); never a benchmark number."""
import argparse, ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg1"); ap.add_argument("--mib", type=float, default=64.0); ap.add_argument("--sstables", type=int, default=0)
ap.add_argument("--repeat", type=int, default=2); ap.add_argument("--host", action="store_true"); ap.add_argument("--rows-per-partition", type=int, default=1000)
a = ap.parse_args()
import torch
from cassandra_b200 import native
wl = dict(bench.WORKLOADS[a.workload]); wl["mib"] = a.mib; wl["rpp"] = a.rows_per_partition
if a.sstables: wl["sstables"] = a.sstables
L = native.lib(); ctx = native.Context(0)
tabs = bench.make_inputs(wl, bench.gpu_compressor(ctx, wl), max(1, (os.cpu_count() or 8)))
u_in = sum(t.compression.data_length for t in tabs); i_in = sum(t.hold[1].numel() for t in tabs); cid, _ = bench.comp_ids(wl)
nout = 1 if not wl["lcs"] else int(2 * u_in // wl["lcs"]) + 4
cap_d = L.b200c_compress_bound(cid, u_in, 16384); cap_i = i_in + (1 << 20); cap_c = u_in // 16384 + 16
if a.host:
    m = bench.build_manifest(tabs, wl); out = bench.OutBufs(nout, cap_d, cap_i, cap_c, device=False)
else:
    dev_in = [(t.hold[0].cuda(), t.hold[1].cuda(), t.hold[2].cuda(), t.summary.cuda()) for t in tabs]
    m = bench.build_manifest(tabs, wl, [(x.data_ptr(), y.data_ptr(), z.data_ptr(), w.data_ptr()) for x, y, z, w in dev_in]); out = bench.OutBufs(nout, cap_d, cap_i, cap_c, device=True)
for r in range(a.repeat):
    res = out.result(); t0 = time.perf_counter()
    ctx.check(L.b200c_compact(ctx.handle, C.byref(m), C.byref(res), 0 if a.host else 1), res.corruption)
    print("run %d: %.1f ms wall, kernels %.1f ms, stages %s, launches %d, MB/s %.0f" % (r, (time.perf_counter() - t0) * 1e3, res.kernel_ms,
          [round(x, 1) for x in ctx.last_stage_ms()], res.kernel_launches, u_in / (time.perf_counter() - t0) / 1e6), flush=True)
ctx.close()
