#!/bin/bash
# 8-GPU round (lean): K10 v1/v2 vs cuBLAS+NCCL, flagship bench at N=8, HPCG at N=8.
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
S=k10bench$$
for r in $(seq 0 $((NG-1))); do timeout 200 python tests/_k10_worker.py --rank $r --world $NG --session $S --device $r --bench > gpurun_out/k10v3_r$r.log 2>&1 & done; wait
tail -2 gpurun_out/k10v3_r0.log | tee gpurun_out/k10v3_bench_n$NG.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29573 bench/coll_sweep.py --min-bytes 1K --max-bytes 16M --step 4 --ops allgather,alltoall,broadcast,reduce_scatter --out gpurun_out/coll_sweep17_small_n$NG.jsonl 2>&1 | grep -v Warning | tail -32 | tee gpurun_out/sweep17_small_n$NG.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $NG --steps 15 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench17_n$NG.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571 recipes/HPCG-Infiniband-IntelMPI/run_hpcg.py --size 256 --seconds 5 2>&1 | tail -2 | tee gpurun_out/hpcg17_n$NG.log
