#!/bin/bash
# round-2 GPU call 3 (2 GPUs): preload shim under torch.distributed (new binding, grouped send/recv, chunked staging), the stock DDP
# ResNet-50 recipe through the CLI with and without the shim, the NVLink litmus, the N-GPU == 1-GPU trainer check;
# plus 1-GPU items that ride along on GPU 0: parity tests (re-calibrated), race-context probe.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NG=$(nvidia-smi -L | wc -l)
export SHIPYARD_TEST_UNVERIFIED=1
timeout 300 python -m pytest tests/test_gpu_coll.py -q -m gpu -x -k "preload or litmus" > gpurun_out/c3_pytest_preload.log 2>&1; tail -5 gpurun_out/c3_pytest_preload.log; tail -12 gpurun_out/preload_test.log
timeout 400 python -m pytest tests/test_gpu_coll.py -q -m gpu -x -k "trainer_equals" > gpurun_out/c3_pytest_trainer.log 2>&1; tail -5 gpurun_out/c3_pytest_trainer.log
unset SHIPYARD_TEST_UNVERIFIED
# stock DDP through the CLI: pool of NG nodes, two jobs (shim / NCCL pass-through)
export SHIPYARD_STATE_DIR=$PWD/gpurun_out/c3_state
rm -rf $SHIPYARD_STATE_DIR; mkdir -p gpurun_out/c3_cfg; cp recipes/PyTorch-GPU/config/stock-ddp/*.yaml gpurun_out/c3_cfg/
sed -i "s/dedicated: 8/dedicated: $NG/" gpurun_out/c3_cfg/pool.yaml
timeout 200 ./shipyard pool add --configdir gpurun_out/c3_cfg -y --raw > gpurun_out/c3_pool.log 2>&1; tail -3 gpurun_out/c3_pool.log
timeout 500 ./shipyard jobs add --configdir gpurun_out/c3_cfg --tail stdout.txt > gpurun_out/c3_jobs.log 2>&1; tail -5 gpurun_out/c3_jobs.log
for j in stockddp-shim stockddp-nccl; do
  d=$SHIPYARD_STATE_DIR/pools/pytorch-gpu/workitems/$j/job-1/train
  echo "== $j"; cat $d/stdout.txt | tail -2; grep -h "shipyard-preload" $d/stderr*.txt | tail -4; tail -3 $d/stderr.txt
  cp $d/stdout.txt gpurun_out/c3_${j}_n$NG.stdout 2>/dev/null; cat $d/stderr*.txt > gpurun_out/c3_${j}_n$NG.stderr 2>/dev/null
done
timeout 100 ./shipyard pool del --configdir gpurun_out/c3_cfg -y > /dev/null 2>&1
rm -rf $SHIPYARD_STATE_DIR
# 1-GPU riders
timeout 400 python -m pytest tests/test_gpu_resnet_parity.py -q -s -m gpu > gpurun_out/c3_parity.log 2>&1; grep -E "^\[|passed|failed|^E  " gpurun_out/c3_parity.log | cut -c1-700
timeout 200 python bench/race_context_probe.py > gpurun_out/c3_race_context.jsonl 2> gpurun_out/c3_race_context.err; cat gpurun_out/c3_race_context.jsonl; tail -3 gpurun_out/c3_race_context.err
