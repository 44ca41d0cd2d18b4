#!/bin/bash
# round-2 GPU call 5 (N GPUs, written for N=2): collective test battery, stock DDP through the CLI (shim vs NCCL), mpiBench --compare sweep
# through the CLI; 1-GPU riders: stager tests + bench, parity tests, bench with the re-race diagnostic.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NG=$(nvidia-smi -L | wc -l)
timeout 700 python -m pytest tests/test_gpu_coll.py -q -m gpu -x > gpurun_out/c5_pytest_coll_n$NG.log 2>&1; tail -6 gpurun_out/c5_pytest_coll_n$NG.log | cut -c1-400
export SHIPYARD_STATE_DIR=$PWD/gpurun_out/c5_state
run_recipe() {   # $1 = recipe config dir, $2 = tag, $3 = pool id
  rm -rf $SHIPYARD_STATE_DIR gpurun_out/c5_cfg; mkdir -p gpurun_out/c5_cfg; cp $1/*.yaml gpurun_out/c5_cfg/
  sed -i "s/dedicated: [0-9]*/dedicated: $NG/" gpurun_out/c5_cfg/pool.yaml
  timeout 200 ./shipyard pool add --configdir gpurun_out/c5_cfg -y --raw > gpurun_out/c5_$2_pool.log 2>&1
  timeout 900 ./shipyard jobs add --configdir gpurun_out/c5_cfg --tail stdout.txt > gpurun_out/c5_$2_jobs.log 2>&1
  for d in $SHIPYARD_STATE_DIR/pools/$3/workitems/*/job-1/*/; do
    j=$(basename $(dirname $(dirname $d)))
    [ -f $d/stdout.txt ] && cp $d/stdout.txt gpurun_out/c5_$2_${j}_n$NG.stdout && cat $d/stderr*.txt > gpurun_out/c5_$2_${j}_n$NG.stderr 2>/dev/null
  done
  timeout 100 ./shipyard pool del --configdir gpurun_out/c5_cfg -y > /dev/null 2>&1
}
run_recipe recipes/PyTorch-GPU/config/stock-ddp ddp pytorch-gpu
for j in stockddp-shim stockddp-nccl; do echo "== $j"; tail -1 gpurun_out/c5_ddp_${j}_n$NG.stdout; grep -h "shipyard-preload" gpurun_out/c5_ddp_${j}_n$NG.stderr | tail -2 | cut -c1-400; done
run_recipe recipes/mpiBench-OpenMPI/config/sweep sweep mpibench
cat gpurun_out/c5_sweep_mpibench-sweep_n$NG.stdout | cut -c1-330; tail -3 gpurun_out/c5_sweep_mpibench-sweep_n$NG.stderr | cut -c1-300
rm -rf $SHIPYARD_STATE_DIR
# ---- 1-GPU riders
timeout 200 python -m pytest tests/test_gpu_stage.py -q -m gpu -x > gpurun_out/c5_stage_tests.log 2>&1; tail -4 gpurun_out/c5_stage_tests.log | cut -c1-300
timeout 200 python bench/stage_bench.py > gpurun_out/c5_stage_bench.json 2> gpurun_out/c5_stage_bench.err; cat gpurun_out/c5_stage_bench.json; tail -2 gpurun_out/c5_stage_bench.err
timeout 400 python -m pytest tests/test_gpu_resnet_parity.py tests/test_zz_gpu_bn_dual.py -q -s -m gpu > gpurun_out/c5_parity.log 2>&1; grep -E "^\[|passed|failed|^E  " gpurun_out/c5_parity.log | cut -c1-400
SHIPYARD_BENCH_RERACE=6 SHIPYARD_CONV_PLAN_DUMP=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-baseline > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; cut -c1-600 gpurun_out/c5_bench.json; grep rerace gpurun_out/c5_bench.err | cut -c1-900
