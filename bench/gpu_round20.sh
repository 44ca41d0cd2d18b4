#!/bin/bash
# 4-GPU round: flagship bench and the plain NCCL/cuDNN arm at N=4 (the one scaling point not measured yet).
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $NG --steps 15 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench20_n$NG.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus $NG --steps 15 --warmup 3 --impl nccl-baseline 2>&1 | tail -1 | tee gpurun_out/base20_n$NG.log
