#!/bin/bash
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu4.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke4.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench4_n1.log
SHIPYARD_NO_TC_GEMM=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench4_n1_notc.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 3500 --csv --log-file gpurun_out/launches_shipyard4.csv \
   python bench.py --gpus 1 --steps 2 --warmup 1 --no-graph --no-e2e > gpurun_out/ncu_ship4.log 2>&1
if [ "$NG" -ge 2 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench4_n$NG.log
fi
