#!/bin/bash
# HPL-MxP on 1 or N GPUs: `gpurun [--gpus N] --timeout 600 -- 'bash bench/gpu_hpl.sh'`: GPU tests of the recipe body, then the benchmark at three sizes
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NG=$(nvidia-smi -L | wc -l)
[ -z "$SKIP_TESTS" ] && timeout 250 python -m pytest tests/test_hpl.py -x -q -m gpu > gpurun_out/h_hpl_tests_n$NG.log 2>&1; tail -3 gpurun_out/h_hpl_tests_n$NG.log | cut -c1-300
if [ "$NG" -gt 1 ]; then L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29591"; else L=python; fi
rm -f gpurun_out/h_hpl_n$NG.log
for cfg in ${HPL_CFGS:-32768:2048 65536:2048 65536:4096}; do set -- ${cfg/:/ }
  timeout 200 $L recipes/HPLinpack-Infiniband-IntelMPI/run_hpl.py -n $1 -b $2 --runs 2 2>&1 | grep -E '^\{|Error' | tail -1 | cut -c1-700 | tee -a gpurun_out/h_hpl_n$NG.log
done
exit 0
