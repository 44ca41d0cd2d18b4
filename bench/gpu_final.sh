#!/bin/bash
# 1-GPU verification round: the driver's own sequence (pytest -m gpu, smoke, bench) plus the time-boxed compute-sanitizer passes
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
T0=$(date +%s)
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/f_pytest_gpu.log 2>&1; tail -4 gpurun_out/f_pytest_gpu.log | cut -c1-300
echo "pytest -m gpu took $(( $(date +%s) - T0 )) s"
timeout 300 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log | cut -c1-300
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench.json'))
cd=d['config']['conv_dispatch']
print('bench: ms', d['ms_per_step'], 'img/s', d['value'], 'e2e', d['e2e']['value'], 'vs eager', d.get('vs_stock_eager'), 'vs tuned', d.get('vs_stock_tuned'))
print({k:v for k,v in cd.items() if k not in ('race_us','halo')})
print('staging', d['e2e'].get('input_staging'))
PY
tail -2 gpurun_out/f_bench.err | cut -c1-300
timeout 200 python recipes/HPCG-Infiniband-IntelMPI/run_hpcg.py --size 256 --seconds 5 2>&1 | tail -1 | cut -c1-900 | tee gpurun_out/f_hpcg_n1.log
timeout 150 python recipes/TensorFlow-Distributed/mnist_replica.py --train_steps 5000 --impl both 2>&1 | grep steps_per_sec | tail -1 | cut -c1-700 | tee gpurun_out/f_tfdist_n1.log
[ -n "$SANITIZE" ] && bash bench/sanitize_timeboxed.sh
[ -n "$COMPILE_BASELINE" ] && SHIPYARD_BASELINE_COMPILE=1 timeout 900 python - <<'PY'
import json, sys
sys.path.insert(0, 'bench')
import stock_baseline
r = stock_baseline.run_compiled(256, 20, 5, 0, 1, 0)
print(json.dumps(r)); open('gpurun_out/f_stock_compiled.json', 'w').write(json.dumps(r))
PY
exit 0
