#!/bin/bash
# round-2 GPU call 21 (1 GPU): K5 on its own stream (streamed pieces) — parity of the streamed paths, then the bench line with the A/B leg (B200C_K5_OVERLAP=0)
mkdir -p gpurun_out
echo "== pytest -m gpu (compaction tests)"; timeout 900 python -m pytest tests/test_gpu_compaction.py -x -q -m gpu 2>&1 | tail -5
echo "== bench cfg1 (+ A/B B200C_K5_OVERLAP=0)"; timeout 1200 python bench.py --ab-env B200C_K5_OVERLAP=0 2>gpurun_out/r21_bench_cfg1.err | tail -1 > gpurun_out/r21_bench_cfg1.json; cut -c1-300 gpurun_out/r21_bench_cfg1.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r21_bench_cfg1.json').read().strip())
print("value",d['value'],"e2e",d['e2e']['value'],"ms",d['ms_per_step'],d['e2e']['ms_per_step'],"verified",d.get('verified',{}).get('ok'))
print(d['roofline'].get('stage_ms')); print(d['e2e'].get('stage_ms')); print({k:v for k,v in d.items() if 'ab' in k.lower()})
PY
tail -5 gpurun_out/r21_bench_cfg1.err | cut -c1-300
