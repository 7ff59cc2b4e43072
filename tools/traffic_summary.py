"""ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv of ONE compaction (tools/one_compaction.py --repeat 1)
-> per-stage DRAM traffic and cold-cache kernel time, as the JSON bench.py reads for roofline.traffic (profiles/r2_traffic.json).
   python tools/traffic_summary.py gpurun_out/r2_traffic_cfg1.csv cfg1 [existing.json] > profiles/r2_traffic.json
Kernels launched before the first k_decompress* / k_index_find* (the input preparation, which compresses the synthetic inputs with K5) are skipped."""
import collections, csv, json, re, sys
STAGE = [("K1 decompress+verify", r"k_decompress|k_lz4_walk|k_lz4_copy"), ("K2 index scan", r"k_index_(find|chain|verify|seq|emit)|k_check_order|k_input_ranges|k_range_plan|k_index_slices"),
         ("K3 partition merge", r"k_merge_|k_bucket_bounds|k_op_first"), ("K4 merge+purge+serialise", r"k_partition_|k_bounds|k_class_hist|k_fanin_scatter|k_tile_|k_sum_stats"),
         ("K4 gather+index", r"k_gather|k_index_simple|k_index_promoted|k_index_sizes|k_add_u64"), ("K5 compress+crc+pack", r"k_compress_chunks|k_lz4_chain_build|k_snappy_chain_build|k_pack_chunks|k_digest|k_offs_add_base"),
         ("meta", r"k_meta_|k_summary_|k_tdrop|k_written_flags"), ("scans", r"k_scan_")]
rows = list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")]))
per = collections.OrderedDict(); started = False; kernels = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]); metric = r["Metric Name"]
    try: v = float(r["Metric Value"].replace(",", ""))
    except ValueError: continue
    unit = r["Metric Unit"]
    if metric.startswith("dram__bytes"): v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    else: v *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1e-6)          # ms
    if re.search(r"k_decompress|k_index_find", name): started = True
    if not started: continue
    stage = next((s for s, pat in STAGE if re.search(pat, name)), "other")
    d = per.setdefault(stage, {"dram_bytes": 0.0, "ms": 0.0}); k = kernels.setdefault(name, {"n": 0, "dram_bytes": 0.0, "ms": 0.0})
    if metric.startswith("dram__bytes"): d["dram_bytes"] += v; k["dram_bytes"] += v
    else: d["ms"] += v; k["ms"] += v; k["n"] += 1
out = {}
if len(sys.argv) > 3:
    try: out = json.load(open(sys.argv[3]))
    except Exception: out = {}
out[sys.argv[2]] = {s: int(d["dram_bytes"]) for s, d in per.items()}
out[sys.argv[2] + "_detail"] = {"per_stage_ms_cold_cache_serialised": {s: round(d["ms"], 2) for s, d in per.items()},
                                "kernels": {k: {"n": v["n"], "dram_GB": round(v["dram_bytes"] / 1e9, 3), "ms": round(v["ms"], 2)} for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])[:24]},
                                "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, one device-resident compaction of the workload"}
json.dump(out, sys.stdout, indent=1)
