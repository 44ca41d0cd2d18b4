#!/bin/bash
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gemm.log
if grep -q "passed" gpurun_out/pytest_gemm.log && ! grep -q "failed" gpurun_out/pytest_gemm.log; then
  timeout 600 python bench/gemm_bench.py 2>&1 | tail -20 | tee gpurun_out/gemm_bench.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 \
   bench/coll_sweep.py --max-bytes 1G --min-bytes 64K --out gpurun_out/coll_sweep_v2_n$NG.jsonl 2>&1 | grep -v Warning | tail -60 | tee gpurun_out/sweep_v2_n$NG.log
for mb in 64 148 256; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29532 \
   bench/coll_sweep.py --min-bytes 16M --max-bytes 256M --ops allreduce,allgather --max-blocks $mb --out gpurun_out/coll_sweep_mb${mb}_n$NG.jsonl 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/sweep_mb${mb}_n$NG.log
done
timeout 600 python -m pytest tests/test_gpu_coll.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_coll3.log
