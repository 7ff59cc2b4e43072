#!/bin/bash
# round-2 GPU call 2: why is the staged K4 slow? launch list + full capture of k_partition_staged at 16 x 64 MiB; parity suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest.log; tail -5 gpurun_out/r2_gputest.log
python tools/one_compaction.py --mib 64 --repeat 2 2>/dev/null | tail -2
B200C_K4_STAGED=0 python tools/one_compaction.py --mib 64 --repeat 2 2>/dev/null | tail -1
B200C_K5=1 python tools/one_compaction.py --mib 64 --repeat 2 2>/dev/null | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_staged.csv python tools/one_compaction.py --mib 64 --repeat 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_staged.csv 2>/dev/null | head -30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_partition_staged -c 1 -o gpurun_out/r2_prof_staged python tools/one_compaction.py --mib 64 --repeat 1 > /dev/null 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_prof_staged.ncu-rep
