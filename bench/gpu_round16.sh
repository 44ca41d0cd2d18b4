#!/bin/bash
# 2-GPU round: HPCG over the halo path, mpiBench recipe through the CLI on device buffers (with stderr), N=2 bench.
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
export SHIPYARD_TEST_QUICK=1
timeout 600 python -m pytest tests/test_gpu_coll.py -m gpu -q -x -k "multi_gpu_collectives or single" 2>&1 | tail -8 | tee gpurun_out/pytest_coll16.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29573 bench/coll_sweep.py --min-bytes 1K --max-bytes 4M --step 4 --ops allgather,alltoall,broadcast --out gpurun_out/coll_sweep16_small_n$NG.jsonl 2>&1 | grep -v Warning | tail -24 | tee gpurun_out/sweep16_small_n$NG.log
S=k10bench$$
for r in $(seq 0 $((NG-1))); do timeout 200 python tests/_k10_worker.py --rank $r --world $NG --session $S --device $r --bench > gpurun_out/k10v3_r$r.log 2>&1 & done; wait
tail -3 gpurun_out/k10v3_r0.log | tee gpurun_out/k10v3_bench_n$NG.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571 recipes/HPCG-Infiniband-IntelMPI/run_hpcg.py --size 256 --seconds 6 2>&1 | tail -2 | tee gpurun_out/hpcg16_n$NG.log
export SHIPYARD_STATE_DIR=$PWD/gpurun_out/state16_$$
sed "s/dedicated: 2/dedicated: $NG/" recipes/mpiBench-OpenMPI/config/pool.yaml > /tmp/pool16.yaml
timeout 200 ./shipyard pool add --configdir recipes/mpiBench-OpenMPI/config --pool /tmp/pool16.yaml -y 2>&1 | tail -3 | tee gpurun_out/recipe16_pool.log
timeout 300 ./shipyard jobs add --configdir recipes/mpiBench-OpenMPI/config --pool /tmp/pool16.yaml --jobs recipes/mpiBench-OpenMPI/config/jobs-gpu.yaml --tail stdout.txt 2>&1 | tail -60 | tee gpurun_out/recipe16_mpibench.log
find $SHIPYARD_STATE_DIR -name "stderr*.txt" -o -name "result.json" | head -5 | while read f; do echo "== $f"; tail -15 $f; done 2>&1 | tee gpurun_out/recipe16_stderr.log
rm -rf $SHIPYARD_STATE_DIR
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $NG --steps 15 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench16_n$NG.log
