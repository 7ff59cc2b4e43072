#!/bin/bash
# round-2 final measurement set (1 GPU): the GPU suite, the bench lines of the three 1-GPU workloads and the reference arm, then the ncu captures
# the roofline numbers come from (launch list + DRAM traffic of one configs[1] compaction, --set full of the top kernels). Outputs: gpurun_out/r2f_*
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
echo "== bench cfg1"; timeout 1500 python bench.py 2>gpurun_out/r2f_bench_cfg1.err | tail -1 > gpurun_out/r2f_bench_cfg1.json; cut -c1-300 gpurun_out/r2f_bench_cfg1.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference 2>/dev/null | tail -1 > gpurun_out/r2f_bench_reference.json; cut -c1-400 gpurun_out/r2f_bench_reference.json
echo "== bench cfg2"; timeout 900 python bench.py --workload cfg2 --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2f_bench_cfg2.json; cut -c1-200 gpurun_out/r2f_bench_cfg2.json
echo "== bench cfg4"; timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2f_bench_cfg4.json; cut -c1-200 gpurun_out/r2f_bench_cfg4.json
echo "== ncu launch list + traffic, one configs[1] compaction"
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2f_traffic_cfg1.csv python tools/one_compaction.py --mib 1024 --repeat 1 > /dev/null 2>&1; echo "traffic rc=$?"
for spec in "k_partition_thr:0:k4_8" "k_partition_thr:1:k4_12" "k_compress_chunks:16:k5" "k_decompress_multi_thr:0:k1"; do
  k=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; tag=${rest#*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c 1 -f -o gpurun_out/r2f_prof_$tag python tools/one_compaction.py --mib 1024 --repeat 1 > /dev/null 2>&1; echo "ncu $tag rc=$?"
  python tools/ncu_top_lines.py gpurun_out/r2f_prof_$tag.ncu-rep 25 > gpurun_out/r2f_top_$tag.txt 2>&1
done
python tools/ncu_summary.py gpurun_out/r2f_prof_k4_8.ncu-rep gpurun_out/r2f_prof_k4_12.ncu-rep gpurun_out/r2f_prof_k5.ncu-rep gpurun_out/r2f_prof_k1.ncu-rep > gpurun_out/r2f_ncu_full_summary.txt 2>&1
rm -f gpurun_out/r2f_prof_*.ncu-rep
ls -la gpurun_out | head -40
