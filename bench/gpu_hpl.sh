#!/bin/bash
# 1-GPU (or N-GPU) call: HPL-MxP tests + benchmark, and the direct-store epilogue numerics in a child process
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NG=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu > gpurun_out/h_gemm_alone.log 2>&1; tail -5 gpurun_out/h_gemm_alone.log | cut -c1-240
timeout 250 python -m pytest tests/test_hpl.py -x -q -m gpu > gpurun_out/h_hpl_tests.log 2>&1; tail -15 gpurun_out/h_hpl_tests.log | cut -c1-300
if [ "$NG" -gt 1 ]; then L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29591"; else L=python; fi
for cfg in "16384 1024" "32768 2048" "65536 2048"; do set -- $cfg
  timeout 200 $L recipes/HPLinpack-Infiniband-IntelMPI/run_hpl.py -n $1 -b $2 --runs 2 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/h_hpl_n$NG.log
done
[ -n "$WITH_BENCH" ] && { timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/h_bench.json'))
    print('bench: ms', d['ms_per_step'], 'img/s', d['value'], 'e2e', d['e2e']['value'], 'vs eager', d.get('vs_stock_eager'), 'vs tuned', d.get('vs_stock_tuned'), 'vs compiled', d.get('vs_stock_compiled'), 'vs_baseline', d.get('vs_baseline'))
    print({k:(v.get('value'), v.get('error')) for k,v in d['baseline_same_run'].items()})
except Exception as e: print('bench parse failed', e)
PY
tail -2 gpurun_out/h_bench.err | cut -c1-300; }
exit 0
