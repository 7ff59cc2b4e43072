#!/bin/bash
# round-2 GPU call 6a (1 GPU): parity after the packed K4 cursor / Snappy in place / adaptive LCS window / progress fix; K4 occupancy; cfg2 + cfg4 lines; CPU arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r6_gputest_default.log 2>&1; echo "pytest(default) rc=$?"; tail -6 gpurun_out/r6_gputest_default.log
echo "== cfg1 256 MiB"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== cfg2"; python tools/one_compaction.py --workload cfg2 --mib 160 --repeat 2 2>/dev/null | tail -1
echo "== cfg2, Snappy chunk copy in shared memory (K5=0)"; B200C_K5=0 python tools/one_compaction.py --workload cfg2 --mib 160 --repeat 2 2>/dev/null | tail -1
timeout 600 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct --clock-control none -k regex:k_partition -c 6 --csv --log-file gpurun_out/r6_k4_occ.csv python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1; echo "ncu k4 rc=$?"
python - <<'PY'
import csv, collections
try:
    rows = [r for r in csv.reader(open("gpurun_out/r6_k4_occ.csv")) if len(r) > 10 and r[0].isdigit()]
    agg = collections.defaultdict(dict)
    for r in rows: agg[(r[0], r[4][:60])][r[-3]] = r[-1]
    for k, d in list(agg.items())[:6]: print(k[1], {a.split("__")[-1][:28]: b for a, b in d.items()})
except Exception as e: print("no csv", e)
PY
timeout 900 python bench.py --workload cfg2 --steps 3 --warmup 3 > gpurun_out/r6_bench_cfg2.json 2> gpurun_out/r6_bench_cfg2.err; echo "cfg2 rc=$?"; tail -1 gpurun_out/r6_bench_cfg2.err | cut -c1-200
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/r6_bench_cfg4.json 2> gpurun_out/r6_bench_cfg4.err; echo "cfg4 rc=$?"; tail -1 gpurun_out/r6_bench_cfg4.err | cut -c1-200
python - <<'PY'
import json
for w in ("cfg2", "cfg4"):
    try:
        d = json.load(open("gpurun_out/r6_bench_%s.json" % w)); print(w, "value", d["value"], "e2e", d["e2e"]["value"], d["roofline"]["stage_ms"], "verified", d.get("verified", {}).get("ok"), "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e: print("no line", w, e)
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r6_ref.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r6_ref.json')); c=d['cpu_baseline']; print('ref', d['value'], c['phase_ms'], c['range_tasks_ms'], c['scaling'])"
