#!/bin/bash
# round-2 GPU call 8 (1 GPU): state after the container was re-created — quick stage clock, bench line without the CPU legs, and
# ncu --set full with source for both K4 thread instantiations (the source-level picture of K4 was never captured)
mkdir -p gpurun_out
echo "== cfg1 256 MiB"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -2
echo "== bench cfg1 (no cpu legs)"; python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r8_bench_cfg1_nocpu.json; cut -c1-600 gpurun_out/r8_bench_cfg1_nocpu.json
for k in "k_partition_thr<8" "k_partition_thr<16"; do
  n=$(echo $k | tr -c 'a-zA-Z0-9\n' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -c 1 -o gpurun_out/r8_prof_$n python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1; echo "ncu $k rc=$?"
  python tools/ncu_top_lines.py gpurun_out/r8_prof_$n.ncu-rep 60 > gpurun_out/r8_top_$n.txt 2>&1
done
ls -la gpurun_out/ | head -20
