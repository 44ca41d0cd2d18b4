#!/bin/bash
# Multi-GPU round for round 2: `gpurun --gpus N --timeout 900 -- 'bash bench/gpu_round2_multi.sh'` with N = 2, 4 or 8 (charged N x).
# Fills the measurement gaps listed in NEXT.md: full-range collective sweep at this N (1 KB - 1 GB where the heap allows), the
# flag-protocol litmus over NVLink (N >= 2), K10, the flagship bench and HPCG, all device-timed, max over ranks.
set -x
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
# 1. litmus (2 ranks are enough) + the multi-GPU test files, quick mode
SHIPYARD_TEST_UNVERIFIED=1 SHIPYARD_TEST_QUICK=1 timeout 600 python -m pytest tests/test_gpu_coll.py -x -q -m gpu > gpurun_out/r2m_pytest_coll_n$NG.log 2>&1; tail -4 gpurun_out/r2m_pytest_coll_n$NG.log
# 2. sweeps: small and large messages, every op, vs NCCL (a 4 GiB heap admits the 1 GB all-reduce; all-gather / all-to-all stop at 256 MB per rank)
SHIPYARD_COLL_HEAP=$((4<<30)) timeout 400 $TR --master-port 29581 bench/coll_sweep.py --min-bytes 1K --max-bytes 1G --step 4 --ops allreduce \
    --out gpurun_out/coll_sweep_r2_allreduce_n$NG.jsonl 2>&1 | grep -v Warning | tail -14 | tee gpurun_out/r2m_sweep_allreduce_n$NG.log
SHIPYARD_COLL_HEAP=$((4<<30)) timeout 400 $TR --master-port 29582 bench/coll_sweep.py --min-bytes 1K --max-bytes 256M --step 4 \
    --ops allgather,alltoall,broadcast,reduce_scatter --out gpurun_out/coll_sweep_r2_others_n$NG.jsonl 2>&1 | grep -v Warning | tail -40 | tee gpurun_out/r2m_sweep_others_n$NG.log
# 3. K10 fused GEMM + all-reduce
S=k10r2$$
for r in $(seq 0 $((NG-1))); do timeout 200 python tests/_k10_worker.py --rank $r --world $NG --session $S --device $r --bench > gpurun_out/r2m_k10_r$r.log 2>&1 & done; wait
tail -2 gpurun_out/r2m_k10_r0.log | tee gpurun_out/r2m_k10_n$NG.log
# 4. flagship bench, HPCG, mpiBench recipe through the CLI
timeout 400 $TR --master-port 29583 bench.py --gpus $NG --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/r2m_bench_n$NG.json
timeout 200 $TR --master-port 29584 recipes/HPCG-Infiniband-IntelMPI/run_hpcg.py --size 256 --seconds 5 2>&1 | tail -2 | tee gpurun_out/r2m_hpcg_n$NG.log
export SHIPYARD_STATE_DIR=$PWD/gpurun_out/r2m_state
timeout 300 ./shipyard pool add --configdir recipes/mpiBench-Infiniband-OpenMPI/config -y --raw > gpurun_out/r2m_recipe_pool.log 2>&1
timeout 300 ./shipyard jobs add --configdir recipes/mpiBench-Infiniband-OpenMPI/config --tail stdout.txt > gpurun_out/r2m_recipe_mpibench_n$NG.log 2>&1; tail -12 gpurun_out/r2m_recipe_mpibench_n$NG.log
rm -rf gpurun_out/r2m_state
