#!/bin/bash
# 2-GPU round: mailbox byte movers + chunked large non-symmetric collectives, small-message sweep, mpiBench recipe on device buffers.
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
export SHIPYARD_TEST_QUICK=1
timeout 600 python -m pytest tests/test_gpu_coll.py -m gpu -q -x -k "multi_gpu_collectives or single" 2>&1 > gpurun_out/pytest_coll18_full.log; tail -8 gpurun_out/pytest_coll18_full.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29573 bench/coll_sweep.py --min-bytes 1K --max-bytes 4M --step 4 --ops allgather,alltoall,broadcast --out gpurun_out/coll_sweep18_small_n$NG.jsonl 2>&1 | grep -v Warning | tail -24 | tee gpurun_out/sweep18_small_n$NG.log
export SHIPYARD_STATE_DIR=$PWD/gpurun_out/state18_$$
sed "s/dedicated: 2/dedicated: $NG/" recipes/mpiBench-OpenMPI/config/pool.yaml > /tmp/pool18.yaml
timeout 200 ./shipyard pool add --configdir recipes/mpiBench-OpenMPI/config --pool /tmp/pool18.yaml -y 2>&1 | tail -3
timeout 300 ./shipyard jobs add --configdir recipes/mpiBench-OpenMPI/config --pool /tmp/pool18.yaml --jobs recipes/mpiBench-OpenMPI/config/jobs-gpu.yaml --tail stdout.txt 2>&1 | tail -90 | tee gpurun_out/recipe18_mpibench.log
rm -rf $SHIPYARD_STATE_DIR
