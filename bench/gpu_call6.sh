#!/bin/bash
# round-2 GPU call 6 (1 GPU): effect of the 256-byte aligned flat parameters on the race and the step; BN dual; forced tc; census
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
SHIPYARD_BENCH_RERACE=4 timeout 400 python bench.py --steps 20 --warmup 5 --no-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c6_bench.json'))
cd=d['config']['conv_dispatch']
print('ms', d['ms_per_step'], 'img/s', d['value'], 'e2e', d['e2e']['value'], {k:v for k,v in cd.items() if k not in ('race_us','halo')})
for k,v in cd['race_us'].items(): print(' ', k, v)
PY
grep rerace gpurun_out/c6_bench.err | cut -c1-700
SHIPYARD_BN_DUAL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-baseline > gpurun_out/c6_bench_dual.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/c6_bench_dual.json')); print('bn_dual ms', d['ms_per_step'], d['value'])"
SHIPYARD_CONV_IMPL=tc timeout 300 python bench.py --steps 20 --warmup 5 --no-baseline > gpurun_out/c6_bench_tc.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/c6_bench_tc.json')); print('forced tc ms', d['ms_per_step'], d['value'])"
SHIPYARD_CONV_IMPL=cudnn timeout 300 python bench.py --steps 20 --warmup 5 --no-baseline > gpurun_out/c6_bench_cudnn.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/c6_bench_cudnn.json')); print('forced cudnn ms', d['ms_per_step'], d['value'])"
timeout 240 python bench/torch_kernel_census.py > gpurun_out/c6_kernel_census.txt 2> gpurun_out/c6_kernel_census.err; head -45 gpurun_out/c6_kernel_census.txt | cut -c1-200
