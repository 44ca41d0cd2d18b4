#!/bin/bash
# Short multi-GPU check of the latency-range collective kernels: `gpurun --gpus 2 --timeout 420 -- 'bash bench/gpu_lm_check.sh'`
#   1. the collective test battery (both transports) + preload shim  2. the mpiBench --compare sweep up to 4 MB through the CLI
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NG=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_gpu_coll.py -q -m gpu -x -k "multi_gpu_collectives or preload" > gpurun_out/lm_pytest_coll_n$NG.log 2>&1; tail -3 gpurun_out/lm_pytest_coll_n$NG.log | cut -c1-400
export SHIPYARD_STATE_DIR=$PWD/gpurun_out/m_state
rm -rf $SHIPYARD_STATE_DIR gpurun_out/m_cfg; mkdir -p gpurun_out/m_cfg; cp recipes/mpiBench-OpenMPI/config/sweep/*.yaml gpurun_out/m_cfg/
sed -i "s/dedicated: [0-9]*/dedicated: $NG/" gpurun_out/m_cfg/pool.yaml
sed -i "s/-e 1G --factor 4/-e 4M --factor 4/" gpurun_out/m_cfg/jobs.yaml
timeout 100 ./shipyard pool add --configdir gpurun_out/m_cfg -y --raw > gpurun_out/lm_pool.log 2>&1
timeout 200 ./shipyard jobs add --configdir gpurun_out/m_cfg --tail stdout.txt > gpurun_out/lm_jobs.log 2>&1
for d in $SHIPYARD_STATE_DIR/pools/mpibench/workitems/*/job-1/*/; do
  [ -f $d/stdout.txt ] && cp $d/stdout.txt gpurun_out/lm_sweep_n$NG.stdout && cat $d/stderr*.txt > gpurun_out/lm_sweep_n$NG.stderr 2>/dev/null
done
timeout 60 ./shipyard pool del --configdir gpurun_out/m_cfg -y > /dev/null 2>&1
python - <<PY
import json
for line in open('gpurun_out/lm_sweep_n$NG.stdout'):
    if line.startswith('{"op"'):
        d = json.loads(line)
        print(f"{d['op']:15s} {d['bytes']:>9d}  sym {d['sy_sym_us']:7.2f}  plain {d['sy_plain_us']:7.2f}  nccl {d['nccl_us']:7.2f}  x{d['speedup_sym_vs_nccl']:.2f}")
PY
tail -2 gpurun_out/lm_sweep_n$NG.stderr | cut -c1-300
rm -rf $SHIPYARD_STATE_DIR
