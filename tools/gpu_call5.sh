#!/bin/bash
# round-2 GPU call 5 (1 GPU): full parity run, K5 distinct-hash fast path A/B, cfg2 (LCS, Snappy) launch list, CPU arm instrumentation
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r5_gputest_default.log 2>&1; echo "pytest(default) rc=$?"; tail -6 gpurun_out/r5_gputest_default.log
echo "== K5 mode 1 (direct)"; B200C_K5=1 python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== K5 mode 2 (direct + distinct-hash fast path)"; B200C_K5=2 python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
B200C_K5=2 timeout 600 python -m pytest tests/test_gpu_codec.py tests/test_gpu_compaction.py -m gpu -q -k "codec or lz4 or golden or synthetic or chunks" > gpurun_out/r5_gputest_k5dup.log 2>&1; echo "pytest(K5=2) rc=$?"; tail -3 gpurun_out/r5_gputest_k5dup.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r5_launches_cfg2.csv python tools/one_compaction.py --workload cfg2 --mib 160 --repeat 1 > gpurun_out/r5_cfg2_run.log 2>&1; echo "ncu cfg2 rc=$?"; tail -1 gpurun_out/r5_cfg2_run.log
python tools/launch_summary.py gpurun_out/r5_launches_cfg2.csv 2>/dev/null | head -40
echo "== cfg2 plain"; python tools/one_compaction.py --workload cfg2 --mib 160 --repeat 2 2>/dev/null | tail -1
for t in 16 64 128; do timeout 600 python bench.py --impl reference --steps 1 --warmup 0 --ref-threads $t > gpurun_out/r5_ref_t$t.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r5_ref_t$t.json')); c=d['cpu_baseline']; print('ref threads $t', d['value'], c['phase_ms'], c['range_tasks_ms'])"; done
