#!/bin/bash
# round-2 closing measurement (1 GPU, final default code): the GPU suite, smoke(), the bench lines of the three 1-GPU workloads and the reference arm
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench cfg1"; timeout 900 python bench.py 2>gpurun_out/r22_bench_cfg1.err | tail -1 > gpurun_out/r22_bench_cfg1.json; cut -c1-200 gpurun_out/r22_bench_cfg1.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference 2>/dev/null | tail -1 > gpurun_out/r22_bench_reference.json; cut -c1-300 gpurun_out/r22_bench_reference.json
echo "== bench cfg2"; timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r22_bench_cfg2.json; cut -c1-200 gpurun_out/r22_bench_cfg2.json
echo "== bench cfg4"; timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r22_bench_cfg4.json; cut -c1-200 gpurun_out/r22_bench_cfg4.json
python - <<'PY'
import json
for n in ("cfg1","cfg2","cfg4"):
    try:
        d=json.loads(open('gpurun_out/r22_bench_%s.json'%n).read().strip())
        print(n,"value",d['value'],"e2e",d['e2e']['value'],"ms",d['ms_per_step'],d['e2e']['ms_per_step'],"verified",d.get('verified',{}).get('ok'),"cpu",d.get('cpu_baseline',{}).get('value'),"frac",d['roofline']['frac'], d['roofline'].get('stage_ms'))
    except Exception as e: print(n,"failed",e)
PY
