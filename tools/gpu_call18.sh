#!/bin/bash
# round-2 GPU call 18 (2 GPUs): ONE compaction of configs[1] sharded by token range over 2 GPUs (value + e2e: host buffers, Index.db streaming inside each shard)
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/r2f_2gpu.err | tail -1 > gpurun_out/r2f_bench_cfg1_2gpu.json
cut -c1-400 gpurun_out/r2f_bench_cfg1_2gpu.json; tail -5 gpurun_out/r2f_2gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_cfg1_2gpu.json').read().strip())
print("value",d['value'],"e2e",d['e2e']['value'],"ms",d['ms_per_step'],d['e2e']['ms_per_step'],"shards",d.get('shards'))
print(d['roofline']['stage_ms']); print(d['e2e']['stage_ms'])
PY
