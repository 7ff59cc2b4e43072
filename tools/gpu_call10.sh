#!/bin/bash
# round-2 GPU call 10 (1 GPU): K5 in two passes (B200C_K5=3) and K1 in two passes (B200C_K1=2) — parity (codec tests + compaction subset incl. the
# corruption tests), A/B against the defaults, launch metrics; then ncu --set full with source for both K4 thread instantiations
mkdir -p gpurun_out
SUB="(test_gpu_codec or golden or synthetic_configs or streaming_matches or corrupt or crc or flipped or config0 or snappy or config2) and not alternate"
echo "== parity with B200C_K5=3"
B200C_K5=3 timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_codec.py tests/test_gpu_compaction.py -k "$SUB" 2>&1 | tail -3
echo "== parity with B200C_K1=2 (batched even for tiny launches)"
B200C_K1=2 B200C_K1_BATCH=2 timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_codec.py tests/test_gpu_compaction.py -k "$SUB" 2>&1 | tail -3
echo "== cfg1 256 MiB: default"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== K5=3"; B200C_K5=3 python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
for b in 4 6 8 12 16; do echo "== K1=2 copy blocks $b"; B200C_K1=2 B200C_K1_COPY_BLOCKS=$b python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1; done
echo "== K1=2 K5=3, 1 GiB"; B200C_K1=2 B200C_K5=3 python tools/one_compaction.py --mib 1024 --repeat 2 2>/dev/null | tail -1
echo "== cfg2 default"; python tools/one_compaction.py --workload cfg2 --mib 160 --repeat 2 2>/dev/null | tail -1; echo "== cfg2 K5=3"; B200C_K5=3 python tools/one_compaction.py --workload cfg2 --mib 160 --repeat 2 2>/dev/null | tail -1
B200C_K5=3 B200C_K1=2 timeout 600 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"chain|k_lz4_" --csv --log-file gpurun_out/r10_k1k5_two_pass.csv python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1; echo "ncu k1k5 rc=$?"
for s in 0 1; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_partition_thr -s $s -c 1 -f -o gpurun_out/r9_prof_k4_$s python tools/one_compaction.py --mib 256 --repeat 1 > gpurun_out/r9_ncu_$s.log 2>&1; echo "ncu k4 #$s rc=$?"
  python tools/ncu_top_lines.py gpurun_out/r9_prof_k4_$s.ncu-rep 70 > gpurun_out/r9_top_k4_$s.txt 2>&1
done
python tools/ncu_summary.py gpurun_out/r9_prof_k4_0.ncu-rep gpurun_out/r9_prof_k4_1.ncu-rep > gpurun_out/r9_ncu_summary.txt 2>&1
rm -f gpurun_out/r9_prof_k4_1.ncu-rep
ls -la gpurun_out/ | head -20
