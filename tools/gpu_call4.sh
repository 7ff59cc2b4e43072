#!/bin/bash
# round-2 GPU call 4: parity (static rows, dev+token-range fix, K4 single-sweep changes), staged K4 convergence A/B, cfg1 bench + reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r4_gputest_default.log 2>&1; echo "pytest(default) rc=$?"; tail -5 gpurun_out/r4_gputest_default.log
B200C_K4_STAGED=1 timeout 900 python -m pytest tests/test_gpu_compaction.py -m gpu -q -k "synthetic or static or streaming or metadata or wide or lcs or golden" > gpurun_out/r4_gputest_staged.log 2>&1; echo "pytest(staged subset) rc=$?"; tail -4 gpurun_out/r4_gputest_staged.log
echo "== default (K4 global)"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== K4 staged"; B200C_K4_STAGED=1 python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
for v in 0 1; do
  B200C_K4_STAGED=$v timeout 600 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct --clock-control none -k regex:k_partition -c 12 --csv --log-file gpurun_out/r4_k4_conv_staged$v.csv python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1; echo "ncu conv staged=$v rc=$?"
done
python - <<'PY'
import csv, collections
for v in (0, 1):
    try:
        rows = [r for r in csv.reader(open("gpurun_out/r4_k4_conv_staged%d.csv" % v)) if len(r) > 10 and r[0].isdigit()]
        agg = collections.defaultdict(dict)
        for r in rows: agg[(r[0], r[4][:60])][r[-3]] = r[-1]
        for k, d in list(agg.items())[:12]: print("staged=%d" % v, k[1], {a.split("__")[-1][:28]: b for a, b in d.items()})
    except Exception as e: print("no conv csv", v, e)
PY
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r4_bench_cfg1.json 2> gpurun_out/r4_bench_cfg1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r4_bench_cfg1.json"))
    print("cfg1 value", d["value"], "e2e", d["e2e"]["value"], "stages", d["roofline"]["stage_ms"]); print("e2e stages", d["e2e"]["stage_ms"]); print("cpu", d.get("cpu_baseline")); print("verified", d.get("verified", {}).get("ok"), d.get("verified", {}).get("problems"))
except Exception as e: print("no bench line", e)
PY
tail -2 gpurun_out/r4_bench_cfg1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4_bench_reference.json 2> gpurun_out/r4_bench_reference.err; echo "reference rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r4_bench_reference.json')); print(d['value'], d['cpu_baseline']['scaling'], d['cpu_baseline']['phase_ms'])"
