#!/bin/bash
# ncu capture (one GPU) of two round-2 kernels inside the real training step (eager, no CUDA graph): the fused all-reduce + SGD kernel
# (world 1: the optimiser pass over the flat parameter buffer) and the 2x2-block max-pool backward.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 170 ncu --set full --clock-control none --import-source on -k regex:'k_fused_sgd_k|k_maxpool_bwd2' -c 2 -o gpurun_out/r2_ncu_sgd_pool -f \
  python bench.py --no-graph --no-baseline --steps 1 --warmup 3 > gpurun_out/r2_ncu_sgd_pool.log 2>&1
tail -3 gpurun_out/r2_ncu_sgd_pool.log | cut -c1-200; ls -la gpurun_out/r2_ncu_sgd_pool.ncu-rep
