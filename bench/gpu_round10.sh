#!/bin/bash
# 1-GPU round: flattened conv wgrad, conv dispatcher (auto / tc / cudnn), clean per-pass conv comparison.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "conv" 2>&1 | tail -12 | tee gpurun_out/pytest_conv10.log
timeout 300 python bench/conv_bench.py 2>&1 | tee gpurun_out/conv_bench10.log
SHIPYARD_CONV_PLAN_DUMP=1 timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench10_n1_auto.log
SHIPYARD_CONV_IMPL=tc timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench10_n1_tc.log
SHIPYARD_CONV_IMPL=cudnn timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench10_n1_cudnn.log
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_ops10.log
