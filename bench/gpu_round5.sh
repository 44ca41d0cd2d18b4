#!/bin/bash
# N-GPU round: correctness of the multi-GPU paths, flagship scaling, collective sweep, K10 bench.
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
export SHIPYARD_TEST_QUICK=1
timeout 600 python -m pytest tests/test_gpu_coll.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu5_n$NG.log
for n in $NG 4; do
  if [ "$n" -le "$NG" ] && [ ! -f gpurun_out/bench5_n$n.log ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 15 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench5_n$n.log
  fi
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $NG --steps 15 --warmup 3 --impl nccl-baseline 2>&1 | tail -1 | tee gpurun_out/base5_n$NG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29562 \
   bench/coll_sweep.py --min-bytes 1K --max-bytes 1G --step 8 --out gpurun_out/coll_sweep5_n$NG.jsonl 2>&1 | grep -v Warning | tail -40 | tee gpurun_out/sweep5_n$NG.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29563 \
   bench/coll_sweep.py --min-bytes 256K --max-bytes 1G --step 16 --ops allgather,reduce_scatter,broadcast --nvls-copy 0 --out gpurun_out/coll_sweep5_p2pcopy_n$NG.jsonl 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/sweep5_p2pcopy_n$NG.log
S=k10bench$$
for r in $(seq 0 $((NG-1))); do timeout 300 python tests/_k10_worker.py --rank $r --world $NG --session $S --device $r --bench > gpurun_out/k10_r$r.log 2>&1 & done; wait
tail -2 gpurun_out/k10_r0.log | tee gpurun_out/k10_bench_n$NG.log
