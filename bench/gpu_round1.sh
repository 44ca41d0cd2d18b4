#!/bin/bash
# First GPU slice: correctness of every native kernel, then short benches.
set -x
export SHIPYARD_COLL_DEBUG=1
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m 2>/dev/null | head -12
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_n1.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --impl nccl-baseline 2>&1 | tail -3 | tee gpurun_out/base_n1.log
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench_n$NG.log
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 10 --warmup 3 --impl nccl-baseline 2>&1 | tail -5 | tee gpurun_out/base_n$NG.log
fi
