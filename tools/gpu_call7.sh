#!/bin/bash
# round-2 GPU call 7 (1 GPU): K1 A/Bs, the measurement set on the BENCH configuration (16 x 1 GiB): launch list + DRAM traffic + ncu --set full of K1, both K4 instantiations and K5
mkdir -p gpurun_out
echo "== cfg1 256 MiB, current build"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
for pad in 16384 40960 90112; do echo "== K1 pad $pad"; B200C_K1_PAD=$pad python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1; done
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r7_traffic_cfg1.csv python tools/one_compaction.py --mib 1024 --repeat 1 > /dev/null 2>&1; echo "traffic rc=$?"
for k in k_decompress_multi_thr "k_partition_thr<8" "k_partition_thr<16" k_compress_chunks_lz4_direct; do
  n=$(echo $k | tr -c 'a-zA-Z0-9\n' '_')
  skip=""; [ "$k" = "k_compress_chunks_lz4_direct" ] && skip="-s 16"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$k" $skip -c 1 -o gpurun_out/r7_prof_$n python tools/one_compaction.py --mib 1024 --repeat 1 > /dev/null 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out/r7_* | head
