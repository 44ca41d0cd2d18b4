#!/bin/bash
# 2-GPU round: the complete GPU suite (multi-GPU tests included) and the TensorFlow-Distributed recipe (fused all-reduce + Adam, CUDA graph).
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
export SHIPYARD_TEST_QUICK=1
( time timeout 1200 python -m pytest tests/ -q -m gpu ) > gpurun_out/pytest_all19.log 2>&1; tail -8 gpurun_out/pytest_all19.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29574 recipes/TensorFlow-Distributed/mnist_replica.py --train_steps 20000 2>&1 | tail -1 | tee gpurun_out/tfdist19_n$NG.log
SHIPYARD_TF_GRAPH=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29575 recipes/TensorFlow-Distributed/mnist_replica.py --train_steps 5000 2>&1 | tail -1 | tee gpurun_out/tfdist19_nograph_n$NG.log
