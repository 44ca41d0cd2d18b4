#!/bin/bash
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu2.log
# launch list of one training step (eager, no graph) with device time per kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 3500 --csv --log-file gpurun_out/launches_shipyard.csv \
   python bench.py --gpus 1 --steps 2 --warmup 1 --no-graph --no-e2e > gpurun_out/ncu_ship.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 \
   bench/coll_sweep.py --max-bytes 1G --out gpurun_out/coll_sweep_n$NG.jsonl 2>&1 | grep -v Warning | tail -80 | tee gpurun_out/sweep_n$NG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29522 \
   bench/coll_sweep.py --max-bytes 256M --dtype bfloat16 --ops allreduce --out gpurun_out/coll_sweep_bf16_n$NG.jsonl 2>&1 | grep -v Warning | tail -30 | tee gpurun_out/sweep_bf16_n$NG.log
