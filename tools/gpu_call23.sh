#!/bin/bash
# round-2 GPU call 23 (1 GPU): the whole GPU suite at the last commit (static counter columns added)
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 300 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -12
