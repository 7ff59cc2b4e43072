#!/bin/bash
# round-2 GPU call 3: parity (staged K4 on / default), A/B, cfg1 bench (+verify, CPU leg), reference arm, traffic + profiles, cfg2 / cfg4 lines
mkdir -p gpurun_out
B200C_K4_STAGED=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3_gputest_staged.log 2>&1; echo "pytest(staged) rc=$?"; tail -6 gpurun_out/r3_gputest_staged.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r3_gputest_default.log 2>&1; echo "pytest(default) rc=$?"; tail -4 gpurun_out/r3_gputest_default.log
echo "== default (K4 global, K5 L1)"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== K4 staged"; B200C_K4_STAGED=1 python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== K4 staged, K5 smem"; B200C_K4_STAGED=1 B200C_K5=0 python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== host path, K4 staged"; B200C_K4_STAGED=1 python tools/one_compaction.py --mib 256 --repeat 3 --host 2>/dev/null | tail -1
B200C_K4_STAGED=1 timeout 1200 python bench.py --steps 3 --warmup 3 --ab-env B200C_K4_STAGED=0 > gpurun_out/r3_bench_cfg1_staged.json 2> gpurun_out/r3_bench_cfg1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r3_bench_cfg1_staged.json"))
    print("cfg1 value", d["value"], "e2e", d["e2e"]["value"], "stages", d["roofline"]["stage_ms"]); print("e2e stages", d["e2e"]["stage_ms"]); print("ab", d["ab"]); print("cpu", d.get("cpu_baseline")); print("verified", d.get("verified"))
except Exception as e: print("no bench line", e)
PY
tail -2 gpurun_out/r3_bench_cfg1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3_bench_reference.json 2> gpurun_out/r3_bench_reference.err; echo "reference rc=$?"; cut -c1-1500 gpurun_out/r3_bench_reference.json
B200C_K4_STAGED=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r3_traffic_cfg1_staged.csv python tools/one_compaction.py --mib 1024 --repeat 1 > /dev/null 2>&1; echo "traffic staged rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r3_traffic_cfg1_global.csv python tools/one_compaction.py --mib 1024 --repeat 1 > /dev/null 2>&1; echo "traffic global rc=$?"
B200C_K4_STAGED=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_partition_staged -c 1 -o gpurun_out/r3_prof_staged python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1; echo "ncu staged rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_compress_chunks_lz4_direct -s 16 -c 1 -o gpurun_out/r3_prof_k5 python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1; echo "ncu k5 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decompress_multi_thr -c 1 -o gpurun_out/r3_prof_k1 python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1; echo "ncu k1 rc=$?"
B200C_K4_STAGED=1 timeout 900 python bench.py --workload cfg2 --steps 3 --warmup 3 > gpurun_out/r3_bench_cfg2.json 2> gpurun_out/r3_bench_cfg2.err; echo "cfg2 rc=$?"; tail -2 gpurun_out/r3_bench_cfg2.err; cut -c1-900 gpurun_out/r3_bench_cfg2.json
B200C_K4_STAGED=1 timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/r3_bench_cfg4.json 2> gpurun_out/r3_bench_cfg4.err; echo "cfg4 rc=$?"; tail -2 gpurun_out/r3_bench_cfg4.err; cut -c1-900 gpurun_out/r3_bench_cfg4.json
ls -la gpurun_out/r3_* | head -20
echo "== schema W 8x256MiB default"; python tools/one_compaction.py --workload cfg4 --mib 256 --repeat 2 2>/dev/null | tail -1
echo "== schema W 8x256MiB wide->warp (32 KiB)"; B200C_K4_WIDE_WARP=32768 python tools/one_compaction.py --workload cfg4 --mib 256 --repeat 2 2>/dev/null | tail -1
B200C_K4_WIDE_WARP=32768 timeout 600 python -m pytest tests/test_gpu_compaction.py -m gpu -q -k "wide or synthetic_configs or streaming_matches or lcs_wide" > gpurun_out/r3_gputest_widewarp.log 2>&1; echo "pytest(wide->warp) rc=$?"; tail -3 gpurun_out/r3_gputest_widewarp.log
