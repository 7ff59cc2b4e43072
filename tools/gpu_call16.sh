#!/bin/bash
# round-2 GPU call 16 (1 GPU): K5 build pass v5 (every random read in the parallel phase)
mkdir -p gpurun_out
B200C_K5=3 timeout 600 python -m pytest -x -q -m gpu tests/test_gpu_codec.py -k "test_gpu_codec" 2>&1 | tail -1
echo "== cfg1 256 MiB default"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== K5=3"; B200C_K5=3 python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== cfg2 K5=3"; B200C_K5=3 python tools/one_compaction.py --workload cfg2 --mib 160 --repeat 2 2>/dev/null | tail -1
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
B200C_K5=3 timeout 600 ncu --metrics $M --clock-control none -k regex:"lz4_chain" -s 32 --csv --log-file gpurun_out/r16_k5.csv python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1
grep -c chain gpurun_out/r16_k5.csv; grep "chain" gpurun_out/r16_k5.csv | cut -d, -f5,10- | cut -c1-200 | tail -12
