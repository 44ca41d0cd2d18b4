#!/bin/bash
# 1-GPU verification round: the driver's own sequence (pytest -m gpu, smoke, bench) plus HPCG / TensorFlow-Distributed at N = 1,
# time-boxed compute-sanitizer passes and (optional, COMPILE_BASELINE=1) the torch.compile flavour of the stock baseline.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
T0=$(date +%s)
timeout 700 python -m pytest tests -q -m gpu > gpurun_out/f_pytest_gpu.log 2>&1; tail -6 gpurun_out/f_pytest_gpu.log | cut -c1-300
grep -E "^(FAILED|ERROR)" gpurun_out/f_pytest_gpu.log | head -20
echo "pytest -m gpu took $(( $(date +%s) - T0 )) s"
timeout 200 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/f_smoke.log | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/f_bench.json'))
    cd=d['config']['conv_dispatch']
    print('bench: ms', d['ms_per_step'], 'img/s', d['value'], 'e2e', d['e2e']['value'], 'vs eager', d.get('vs_stock_eager'), 'vs tuned', d.get('vs_stock_tuned'), 'vs_baseline', d.get('vs_baseline'))
    print({k:v for k,v in cd.items() if k not in ('race_us','halo')})
    print('staging', d['e2e'].get('input_staging'), 'launches', d.get('gpu_launches'), 'clocks', d.get('clocks'))
except Exception as e: print('bench parse failed', e)
PY
tail -2 gpurun_out/f_bench.err | cut -c1-300
timeout 150 python recipes/HPCG-Infiniband-IntelMPI/run_hpcg.py --size 256 --seconds 5 2>&1 | tail -1 | cut -c1-900 | tee gpurun_out/f_hpcg_n1.log
timeout 120 python recipes/TensorFlow-Distributed/mnist_replica.py --train_steps 5000 --impl both 2>&1 | grep steps_per_sec | tail -1 | cut -c1-700 | tee gpurun_out/f_tfdist_n1.log
echo "elapsed $(( $(date +%s) - T0 )) s before sanitizers"
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  timeout ${SAN_TIMEOUT:-80} $CS --tool $tool --error-exitcode 9 python bench/sanitize_quick.py > gpurun_out/sanitize_quick_$tool.log 2>&1; echo "== $tool rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|kernels ran' gpurun_out/sanitize_quick_$tool.log | tr '\n' ' ' | cut -c1-300)"
done
echo "elapsed $(( $(date +%s) - T0 )) s before compile baseline"
[ -n "$COMPILE_BASELINE" ] && SHIPYARD_BASELINE_COMPILE=1 timeout 240 python - <<'PY'
import json, sys
sys.path.insert(0, 'bench')
import stock_baseline
try:
    r = stock_baseline.run_compiled(256, 20, 5, 0, 1, 0)
    print(json.dumps(r)[:600]); open('gpurun_out/f_stock_compiled.json', 'w').write(json.dumps(r))
except Exception as e:
    print('compiled baseline failed:', repr(e)[:500])
PY
echo "total $(( $(date +%s) - T0 )) s"
exit 0
