#!/bin/bash
# 1-GPU round: fixed wgrad finalize, BN grid sweep, bench with/without the tcgen05 GEMM path, graph-captured HPCG.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_hpcg.py -m gpu -x -q -k "nt_wgrad or existing_grad or autograd or hpcg_gpu_single" 2>&1 | tail -15 | tee gpurun_out/pytest_gemm7.log
timeout 200 python bench/wgrad_bench.py 2>&1 | tee gpurun_out/wgrad_bench7.log
for b in 5,3,2 5,2,2 5,4,2 8,3,2 5,3,4; do timeout 100 python bench/bn_bench.py --big-only --bps3 $b --iters 12 2>&1 | tail -1 | tee -a gpurun_out/bn_bench7.log; done
timeout 100 python bench/bn_bench.py --big-only --iters 12 2>&1 | tee -a gpurun_out/bn_bench7.log
timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench7_n1.log
SHIPYARD_NO_TC_GEMM=1 timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench7_n1_notc.log
timeout 200 python recipes/HPCG-Infiniband-IntelMPI/run_hpcg.py --n 128 --t 4 2>&1 | tail -2 | tee gpurun_out/hpcg7_n1.log
timeout 200 python recipes/HPCG-Infiniband-IntelMPI/run_hpcg.py --n 256 --t 6 2>&1 | tail -2 | tee -a gpurun_out/hpcg7_n1.log
