#!/bin/bash
# 1-GPU round: conv tests (fprop/dgrad/wgrad), end-to-end bench in three kernel configurations, per-kernel attribution for each.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "conv" 2>&1 | tail -15 | tee gpurun_out/pytest_conv9.log
SHIPYARD_TC_CONV=1 timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench9_n1_tcconv.log
SHIPYARD_TC_CONV=0 timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench9_n1_tcgemm.log
SHIPYARD_TC_CONV=0 SHIPYARD_NO_TC_GEMM=1 timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench9_n1_notc.log
for cfg in "1 0" "0 1"; do set -- $cfg
  tag=conv$1_notc$2
  SHIPYARD_TC_CONV=$1 SHIPYARD_NO_TC_GEMM=$( [ "$2" = 1 ] && echo 1 ) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 2400 --csv \
     --log-file gpurun_out/launches9_$tag.csv python bench.py --gpus 1 --steps 2 --warmup 1 --no-graph --no-e2e > gpurun_out/ncu9_$tag.log 2>&1
done
