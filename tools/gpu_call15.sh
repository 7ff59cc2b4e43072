#!/bin/bash
# round-2 GPU call 15 (1 GPU): K5 build pass v4 (table-only shared memory, 14 chunks per SM, bytes fetched four steps ahead); K4 diet 2 reverted
mkdir -p gpurun_out
echo "== parity with B200C_K5=3"
B200C_K5=3 timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_codec.py tests/test_gpu_compaction.py -k "(test_gpu_codec or golden or synthetic_configs or streaming_matches or snappy or config2) and not alternate" 2>&1 | tail -2
echo "== parity default (K4)"; timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_compaction.py -k "golden or synthetic_configs or edge or mixed or static or range_tombstone or scratch_overflow or metadata" 2>&1 | tail -2
echo "== cfg1 256 MiB default"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== K5=3"; B200C_K5=3 python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== cfg2 default"; python tools/one_compaction.py --workload cfg2 --mib 160 --repeat 2 2>/dev/null | tail -1
echo "== cfg2 K5=3"; B200C_K5=3 python tools/one_compaction.py --workload cfg2 --mib 160 --repeat 2 2>/dev/null | tail -1
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,dram__bytes_read.sum,dram__bytes_write.sum
B200C_K5=3 timeout 600 ncu --metrics $M --clock-control none -k regex:"chain" -s 32 --csv --log-file gpurun_out/r15_k5.csv python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r15_k5.csv'))); hdr=None; per={}
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)<len(hdr): continue
    d=dict(zip(hdr,r)); per.setdefault((int(d['ID']),d['Kernel Name'][:34],d['Grid Size']),{})[d['Metric Name'].replace('smsp__average_warps_issue_stalled_','st_').replace('_per_issue_active.ratio','').split('.')[0][-24:]]=d['Metric Value']
for k in sorted(per): print(k, per[k])
PY
