#!/bin/bash
# round-2 GPU call 9 (1 GPU): ncu --set full with source for both K4 thread instantiations (call 8's -k regex carried template arguments,
# which the default kernel-name base does not contain: nothing was captured)
mkdir -p gpurun_out
for s in 0 1; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_partition_thr -s $s -c 1 -f -o gpurun_out/r9_prof_k4_$s python tools/one_compaction.py --mib 256 --repeat 1 > gpurun_out/r9_ncu_$s.log 2>&1; echo "ncu k4 #$s rc=$?"
  python tools/ncu_top_lines.py gpurun_out/r9_prof_k4_$s.ncu-rep 70 > gpurun_out/r9_top_k4_$s.txt 2>&1
done
python tools/ncu_summary.py gpurun_out/r9_prof_k4_0.ncu-rep gpurun_out/r9_prof_k4_1.ncu-rep > gpurun_out/r9_ncu_summary.txt 2>&1
ls -la gpurun_out/ | head -20
