#!/bin/bash
# 1-GPU round: CTA-pair (cta_group::2) GEMM / conv kernels, reordered split-K wgrad.
set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "two_cta" 2>&1 | tail -15 | tee gpurun_out/pytest_2cta12.log
timeout 240 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "conv_implicit or wgrad" 2>&1 | tail -5 | tee gpurun_out/pytest_wgrad12.log
timeout 200 python bench/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench12.log
timeout 300 python bench/conv_bench.py 2>&1 | tee gpurun_out/conv_bench12.log
timeout 200 python bench/wgrad_bench.py 2>&1 | tee gpurun_out/wgrad_bench12.log
