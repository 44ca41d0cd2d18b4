#!/bin/bash
# Multi-GPU round: `gpurun --gpus N --timeout 1800 -- 'bash bench/gpu_multi.sh'` with N = 2, 4 or 8 (charged N x).
#   1. collective test battery (all ranks, both transports; quick mode at N = 8)
#   2. north-star config #2 literally: stock DDP ResNet-50 through the CLI with the preload shim, then with SHIPYARD_COLL_DISABLE=1
#   3. north-star config #4: mpiBench --compare sweep 1 KB - 1 GB through the CLI (shipyard kernels vs C-level NCCL)
#   4. flagship bench (same-run stock baselines + collective block), strong-scaling point (global batch 256)
#   5. HPCG (validity gate + GFLOP/s), TensorFlow-Distributed --impl both, K10
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
Q=""; [ "$NG" -ge 8 ] && Q=1
SHIPYARD_TEST_QUICK=$Q timeout 900 python -m pytest tests/test_gpu_coll.py -q -m gpu -x ${Q:+-k "multi_gpu_collectives and auto or preload or litmus"} > gpurun_out/m_pytest_coll_n$NG.log 2>&1; tail -3 gpurun_out/m_pytest_coll_n$NG.log | cut -c1-300
export SHIPYARD_STATE_DIR=$PWD/gpurun_out/m_state
run_recipe() {   # $1 = recipe config dir, $2 = tag, $3 = pool id
  rm -rf $SHIPYARD_STATE_DIR gpurun_out/m_cfg; mkdir -p gpurun_out/m_cfg; cp $1/*.yaml gpurun_out/m_cfg/
  sed -i "s/dedicated: [0-9]*/dedicated: $NG/" gpurun_out/m_cfg/pool.yaml
  [ -n "$4" ] && sed -i "$4" gpurun_out/m_cfg/jobs.yaml
  timeout 200 ./shipyard pool add --configdir gpurun_out/m_cfg -y --raw > gpurun_out/m_$2_pool.log 2>&1
  timeout 900 ./shipyard jobs add --configdir gpurun_out/m_cfg --tail stdout.txt > gpurun_out/m_$2_jobs.log 2>&1
  for d in $SHIPYARD_STATE_DIR/pools/$3/workitems/*/job-1/*/; do
    j=$(basename $(dirname $(dirname $d)))
    [ -f $d/stdout.txt ] && cp $d/stdout.txt gpurun_out/m_$2_${j}_n$NG.stdout && cat $d/stderr*.txt > gpurun_out/m_$2_${j}_n$NG.stderr 2>/dev/null
  done
  timeout 100 ./shipyard pool del --configdir gpurun_out/m_cfg -y > /dev/null 2>&1
}
run_recipe recipes/PyTorch-GPU/config/stock-ddp ddp pytorch-gpu "s/--steps 20 --warmup 5/--steps 15 --warmup 4/"
for j in stockddp-shim stockddp-nccl; do echo "== $j"; tail -1 gpurun_out/m_ddp_${j}_n$NG.stdout | cut -c1-420; grep -h "collectives on shipyard" gpurun_out/m_ddp_${j}_n$NG.stderr | tail -1 | cut -c1-300; done
run_recipe recipes/mpiBench-OpenMPI/config/sweep sweep mpibench
cat gpurun_out/m_sweep_mpibench-sweep_n$NG.stdout | cut -c1-330; tail -2 gpurun_out/m_sweep_mpibench-sweep_n$NG.stderr | cut -c1-300
rm -rf $SHIPYARD_STATE_DIR
timeout 600 $TR --master-port 29583 bench.py --gpus $NG --steps 15 --warmup 5 2> gpurun_out/m_bench_n$NG.err | tail -1 > gpurun_out/m_bench_n$NG.json; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/m_bench_n$NG.json'))
    print('bench N=$NG', d['value'], 'img/s', d['ms_per_step'], 'ms; e2e', d['e2e']['value'], '; vs eager', d.get('vs_stock_eager'), 'vs tuned', d.get('vs_stock_tuned'))
    print('  baselines', {k:(v.get('value'), v.get('e2e',{}).get('value'), v.get('what','')[-60:]) for k,v in d.get('baseline_same_run',{}).items()})
    print('  collectives', json.dumps(d.get('collectives'))[:900])
except Exception as e: print('bench parse failed', e)
PY
tail -3 gpurun_out/m_bench_n$NG.err | cut -c1-300
B=$((256 / NG)); SHIPYARD_BENCH_BATCH=$B timeout 400 $TR --master-port 29585 bench.py --gpus $NG --steps 15 --warmup 5 --no-baseline --no-coll 2>/dev/null | tail -1 > gpurun_out/m_bench_strong_n$NG.json; python -c "
import json; d=json.load(open('gpurun_out/m_bench_strong_n$NG.json')); print('strong scaling (global batch 256): N=$NG', d['value'], 'img/s', d['ms_per_step'], 'ms')" 2>&1 | tail -1
timeout 300 $TR --master-port 29584 recipes/HPCG-Infiniband-IntelMPI/run_hpcg.py --size 256 --seconds 5 2>&1 | tail -2 | cut -c1-900 | tee gpurun_out/m_hpcg_n$NG.log
timeout 200 $TR --master-port 29586 recipes/TensorFlow-Distributed/mnist_replica.py --train_steps 5000 --impl both 2>&1 | grep steps_per_sec | tail -1 | cut -c1-700 | tee gpurun_out/m_tfdist_n$NG.log
S=k10m$$
for r in $(seq 0 $((NG-1))); do timeout 200 python tests/_k10_worker.py --rank $r --world $NG --session $S --device $r --bench > gpurun_out/m_k10_r$r.log 2>&1 & done; wait
tail -1 gpurun_out/m_k10_r0.log | cut -c1-400 | tee gpurun_out/m_k10_n$NG.log
