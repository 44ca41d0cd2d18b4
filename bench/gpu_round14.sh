#!/bin/bash
# 1-GPU round: column-sticky stats walk, 2-CTA dgrad variants, dispatcher with tc2 candidates + graph timing.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 > gpurun_out/pytest_gemm14_full.log; tail -5 gpurun_out/pytest_gemm14_full.log
timeout 200 python bench/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench14.log | tail -3
timeout 300 python bench/conv_bench.py 2>&1 | tee gpurun_out/conv_bench14.log | tail -3
SHIPYARD_CONV_PLAN_DUMP=1 timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench14_n1_auto.log
SHIPYARD_CONV_IMPL=tc timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench14_n1_tc.log
