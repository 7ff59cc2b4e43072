#!/bin/bash
# round-2 GPU call 12 (1 GPU): Index.db streaming (K2 per piece, pieces planned on the host) — the whole GPU suite, then the bench line of
# configs[1] without the CPU legs (value + e2e), e2e with one piece for comparison
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
echo "== bench cfg1 (no cpu legs)"; python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-ranges 1 2>/dev/null | tail -1 > gpurun_out/r12_bench_cfg1_nocpu.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r12_bench_cfg1_nocpu.json').read().strip())
print("value",d['value'],"e2e",d['e2e']['value'],"e2e ms",d['e2e']['ms_per_step'])
print("stage_ms",d['roofline']['stage_ms']); print("e2e stage_ms",d['e2e']['stage_ms']); print("sweep",d.get('e2e_ranges_sweep'))
PY
echo "== cfg1 256 MiB default"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
echo "== K5=3 (build pass v2)"; B200C_K5=3 python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
B200C_K5=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_compress_chunks_lz4_chain -s 16 -c 1 -f -o gpurun_out/r12_prof_k5b python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1; echo "ncu k5b rc=$?"
python tools/ncu_top_lines.py gpurun_out/r12_prof_k5b.ncu-rep 45 > gpurun_out/r12_top_k5b.txt 2>&1
B200C_K5=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lz4_chain_build -s 16 -c 1 -f -o gpurun_out/r12_prof_k5a python tools/one_compaction.py --mib 256 --repeat 1 > /dev/null 2>&1; echo "ncu k5a rc=$?"
python tools/ncu_top_lines.py gpurun_out/r12_prof_k5a.ncu-rep 30 > gpurun_out/r12_top_k5a.txt 2>&1
python tools/ncu_summary.py gpurun_out/r12_prof_k5a.ncu-rep gpurun_out/r12_prof_k5b.ncu-rep > gpurun_out/r12_ncu_summary.txt 2>&1
rm -f gpurun_out/r12_prof_k5a.ncu-rep
ls -la gpurun_out | head -30
