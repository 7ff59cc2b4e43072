#!/bin/bash
# round-2 GPU call 20 (1 GPU): the GPU suite with the counter-column tests
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15
