#!/bin/bash
# round-2 GPU call 14 (1 GPU): K4 code diet round 2 (Sink helpers, stats and small loops out of line), equal-piece schedule of the host pipeline
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
echo "== cfg1 256 MiB default"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== cfg4 (schema W) 256 MiB"; python tools/one_compaction.py --workload cfg4 --mib 256 --repeat 2 2>/dev/null | tail -1
echo "== bench cfg1 (no cpu legs), A/B geometric schedule, sweep of equal pieces"; python bench.py --steps 3 --warmup 3 --no-cpu-baseline --ab-env B200C_SCHEDULE=geometric --e2e-ranges 6,12,16 2>/dev/null | tail -1 > gpurun_out/r14_bench_cfg1_nocpu.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r14_bench_cfg1_nocpu.json').read().strip())
print("value",d['value'],"e2e",d['e2e']['value'],"e2e ms",d['e2e']['ms_per_step'])
print("stage_ms",d['roofline']['stage_ms']); print("e2e stage_ms",d['e2e']['stage_ms']); print("ab",d.get('ab')); print("sweep",d.get('e2e_ranges_sweep'))
PY
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
timeout 600 ncu --metrics $M --clock-control none -k regex:"k_partition_thr" -c 3 --csv --log-file gpurun_out/r14_k4.csv python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r14_k4.csv'))); hdr=None; per={}
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)<len(hdr): continue
    d=dict(zip(hdr,r)); per.setdefault((int(d['ID']),d['Kernel Name'][:34],d['Grid Size']),{})[d['Metric Name'].replace('smsp__average_warps_issue_stalled_','st_').replace('_per_issue_active.ratio','').split('.')[0][-24:]]=d['Metric Value']
for k in sorted(per): print(k, per[k])
PY
