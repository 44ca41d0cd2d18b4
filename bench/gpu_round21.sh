#!/bin/bash
# final 1-GPU sanity: what the driver runs at round end.
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/pytest_all21.log 2>&1; tail -4 gpurun_out/pytest_all21.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke21.log
timeout 200 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench21_default.log
