#!/bin/bash
# round-2 GPU call 6 (2 GPUs): ONE compaction sharded by token range over two ranks, as the driver launches it; then the reference arm the same way
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r6_bench_2gpu.json 2> gpurun_out/r6_bench_2gpu.err; echo "bench 2gpu rc=$?"
tail -5 gpurun_out/r6_bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r6_bench_2gpu.json").read().strip().splitlines()[-1])
    print("2gpu value", d["value"], "e2e", d["e2e"]["value"], "scaling", d["scaling"], "ms", d["ms_per_step"]); print({k: d.get(k) for k in ("verified", "shards", "config")})
except Exception as e: print("no 2gpu line", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r6_bench_reference_2gpu.json 2> gpurun_out/r6_bench_reference_2gpu.err; echo "reference 2gpu rc=$?"; tail -2 gpurun_out/r6_bench_reference_2gpu.err | cut -c1-300; cut -c1-400 gpurun_out/r6_bench_reference_2gpu.json
