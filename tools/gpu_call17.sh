#!/bin/bash
# round-2 GPU call 17 (1 GPU): multi-cell (complex) columns on the GPU + the whole suite (the K4 kernels gained a CX instantiation)
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
echo "== cfg1 256 MiB default"; python tools/one_compaction.py --mib 256 --repeat 3 2>/dev/null | tail -1
