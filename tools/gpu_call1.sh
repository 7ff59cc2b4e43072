#!/bin/bash
# round-2 GPU call 1: parity suite, then K4 staged vs global A/B at 16 x 256 MiB, then the configs[1] bench line
mkdir -p gpurun_out
( nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; which java javac ) > gpurun_out/r2_box.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest.log
tail -15 gpurun_out/r2_gputest.log
timeout 600 python bench.py --sstable-mib 256 --steps 3 --warmup 3 --no-cpu-baseline --ab-env B200C_K4_STAGED=0 > gpurun_out/r2_bench_256_ab.json 2> gpurun_out/r2_bench_256_ab.err; echo "bench256 rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_256_ab.json"))
    print("256MiB value", d["value"], "e2e", d["e2e"]["value"], "stages", d["roofline"]["stage_ms"], "AB(global K4)", d["ab"])
except Exception as e: print("no bench line", e)
PY
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_1g.json 2> gpurun_out/r2_bench_1g.err; echo "bench1g rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_1g.json"))
    print("1GiB value", d["value"], "e2e", d["e2e"]["value"], "stages", d["roofline"]["stage_ms"], "e2e stages", d["e2e"]["stage_ms"])
except Exception as e: print("no bench line", e)
PY
tail -5 gpurun_out/r2_bench_1g.err
