#!/bin/bash
# 1-GPU (or N-GPU) call: HPL-MxP tests + benchmark, and the direct-store epilogue numerics in a child process
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NG=$(nvidia-smi -L | wc -l)
SHIPYARD_GEMM_DIRECT_STORE=1 timeout 200 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu -k "matches_fp32 or bias_stats or two_cta or conv_implicit or nn_mn_major or conv1x1_and_linear" > gpurun_out/h_direct.log 2>&1; tail -40 gpurun_out/h_direct.log | cut -c1-240
timeout 250 python -m pytest tests/test_hpl.py -x -q -m gpu > gpurun_out/h_hpl_tests.log 2>&1; tail -15 gpurun_out/h_hpl_tests.log | cut -c1-300
if [ "$NG" -gt 1 ]; then L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29591"; else L=python; fi
for cfg in "16384 1024" "32768 2048" "65536 2048"; do set -- $cfg
  timeout 200 $L recipes/HPLinpack-Infiniband-IntelMPI/run_hpl.py -n $1 -b $2 --runs 2 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/h_hpl_n$NG.log
done
