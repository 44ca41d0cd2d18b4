#!/bin/bash
# 1-GPU call: new stem kernels (tests, bench, ncu), re-calibrated parity tests, bench with same-run baselines
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_stem.py -q -m gpu -x > gpurun_out/c4_stem_tests.log 2>&1; tail -15 gpurun_out/c4_stem_tests.log | cut -c1-300
timeout 120 python bench/stem_bench.py > gpurun_out/c4_stem_bench.json 2> gpurun_out/c4_stem_bench.err; cat gpurun_out/c4_stem_bench.json; tail -3 gpurun_out/c4_stem_bench.err
timeout 400 python -m pytest tests/test_gpu_resnet_parity.py -q -s -m gpu > gpurun_out/c4_parity.log 2>&1; grep -E "^\[|passed|failed|^E  " gpurun_out/c4_parity.log | cut -c1-500
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; cat gpurun_out/c4_bench.json; tail -5 gpurun_out/c4_bench.err
SHIPYARD_STEM_IMPL=cudnn timeout 300 python bench.py --steps 20 --warmup 5 --no-baseline > gpurun_out/c4_bench_stem_cudnn.json 2> gpurun_out/c4_bench_stem_cudnn.err; cat gpurun_out/c4_bench_stem_cudnn.json | cut -c1-400
timeout 200 python bench/race_context_probe.py 2> gpurun_out/c4_race.err | grep "G weights" ; tail -2 gpurun_out/c4_race.err
for v in fprop wgrad; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:stem_s2d -s 1 -c 1 -f -o gpurun_out/c4_ncu_stem_$v python bench/stem_bench.py --one $v > gpurun_out/c4_ncu_stem_$v.log 2>&1; tail -1 gpurun_out/c4_ncu_stem_$v.log
done
