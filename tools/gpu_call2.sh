#!/bin/bash
# round-2 GPU call: parity suite, K4 staged vs global / K5 smem vs L1 A/B, the configs[1] bench line with verification + CPU arm, then profiles
mkdir -p gpurun_out
( nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; which java javac ) > gpurun_out/r2_box.txt 2>&1
B200C_K4_STAGED=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest.log; tail -12 gpurun_out/r2_gputest.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gputest_default.log 2>&1; echo "pytest(default switches) rc=$?"; tail -3 gpurun_out/r2_gputest_default.log
echo "== default (K4 global)"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== K4 staged"; B200C_K4_STAGED=1 python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== K5 L1"; B200C_K5=1 python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== host path staged"; B200C_K4_STAGED=1 python tools/one_compaction.py --mib 256 --repeat 3 --host 2>/dev/null | tail -1
B200C_K4_STAGED=1 timeout 1200 python bench.py --steps 3 --warmup 3 --ab-env B200C_K4_STAGED=0 > gpurun_out/r2_bench_cfg1.json 2> gpurun_out/r2_bench_cfg1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_cfg1.json"))
    print("cfg1 value", d["value"], "e2e", d["e2e"]["value"], "stages", d["roofline"]["stage_ms"]); print("e2e stages", d["e2e"]["stage_ms"]); print("cpu", d.get("cpu_baseline")); print("verified", d.get("verified"))
except Exception as e: print("no bench line", e)
PY
tail -3 gpurun_out/r2_bench_cfg1.err
B200C_K4_STAGED=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python tools/one_compaction.py --mib 64 --repeat 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_launches.csv 2>/dev/null | head -24
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_partition_staged -c 1 -o gpurun_out/r2_prof_staged env B200C_K4_STAGED=1 python tools/one_compaction.py --mib 64 --repeat 1 > /dev/null 2>&1; echo "ncu staged rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_compress_chunks -s 20 -c 1 -o gpurun_out/r2_prof_k5 python tools/one_compaction.py --mib 64 --repeat 1 > /dev/null 2>&1; echo "ncu k5 rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
