#!/bin/bash
# round-2 GPU call 19 (1 GPU): K2 by Summary intervals (count + prove, scan, emit) — the GPU suite, the stage clock with and without it, the bench line
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
echo "== pytest parity, legacy K2"; B200C_K2_LEGACY=1 timeout 600 python -m pytest tests/test_gpu_compaction.py -x -q -m gpu 2>&1 | tail -2
echo "== 256 MiB x 16, default"; timeout 300 python tools/one_compaction.py --mib 256 --repeat 3 2>&1 | tail -1
echo "== 256 MiB x 16, legacy K2"; B200C_K2_LEGACY=1 timeout 300 python tools/one_compaction.py --mib 256 --repeat 3 2>&1 | tail -1
echo "== bench cfg1 (+ A/B legacy K2)"; timeout 1200 python bench.py --ab-env B200C_K2_LEGACY=1 2>gpurun_out/r19_bench_cfg1.err | tail -1 > gpurun_out/r19_bench_cfg1.json; cut -c1-300 gpurun_out/r19_bench_cfg1.json
grep -i "ab-env\|A/B\|legacy" gpurun_out/r19_bench_cfg1.err | tail -8
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r19_bench_cfg1.json').read().strip())
print("value",d['value'],"e2e",d['e2e']['value'],"ms",d['ms_per_step'],d['e2e']['ms_per_step'],"verified",d.get('verified'))
print(d['roofline'].get('stage_ms')); print(d['e2e'].get('stage_ms')); print({k:v for k,v in d.items() if 'ab' in k.lower()})
PY
