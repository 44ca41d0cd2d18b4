#!/bin/bash
# 2-GPU round: K10 v2 (GEMM + reduce-scatter/all-gather), NCCL preload shim, HPCG halo path, collectives regression, N=2 bench.
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
export SHIPYARD_TEST_QUICK=1
timeout 900 python -m pytest tests/test_gpu_coll.py tests/test_hpcg.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_multi11.log
S=k10bench$$
for r in $(seq 0 $((NG-1))); do timeout 300 python tests/_k10_worker.py --rank $r --world $NG --session $S --device $r --bench > gpurun_out/k10v2_r$r.log 2>&1 & done; wait
tail -2 gpurun_out/k10v2_r0.log | tee gpurun_out/k10v2_bench_n$NG.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571 recipes/HPCG-Infiniband-IntelMPI/run_hpcg.py --n 256 --t 6 2>&1 | tail -2 | tee gpurun_out/hpcg11_n$NG.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $NG --steps 15 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench11_n$NG.log
# whole stack on GPUs: pool add (probe, node prep, cascade) + jobs add (native task runner -> MPI face -> device kernels)
export SHIPYARD_STATE_DIR=$PWD/gpurun_out/state11_$$
sed "s/dedicated: 2/dedicated: $NG/" recipes/mpiBench-OpenMPI/config/pool.yaml > /tmp/pool11.yaml
timeout 200 ./shipyard pool add --configdir recipes/mpiBench-OpenMPI/config --pool /tmp/pool11.yaml -y 2>&1 | tail -5 | tee gpurun_out/recipe11_pool.log
timeout 300 ./shipyard jobs add --configdir recipes/mpiBench-OpenMPI/config --pool /tmp/pool11.yaml --jobs recipes/mpiBench-OpenMPI/config/jobs-gpu.yaml --tail stdout.txt 2>&1 | tail -45 | tee gpurun_out/recipe11_mpibench.log
rm -rf $SHIPYARD_STATE_DIR
