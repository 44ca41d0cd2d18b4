#!/bin/bash
# 1-GPU round: two-group pipelined GEMM epilogue, TMA-im2col implicit-GEMM conv (fprop/dgrad), HPCG graph test.
set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_gemm.py tests/test_hpcg.py -m gpu -x -q -k "not multi" 2>&1 | tail -15 | tee gpurun_out/pytest_gemm8.log
timeout 200 python bench/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench8.log
timeout 200 python bench/conv_bench.py 2>&1 | tee gpurun_out/conv_bench8.log
timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench8_n1.log
