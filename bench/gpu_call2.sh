#!/bin/bash
# round-2 GPU call 2 (1 GPU): full-model parity tests, 1x1 dgrad probe, TMA overlap probe, ncu of the NN / TN GEMM on the 64x64 layer
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 60 python bench/tma_overlap_probe.py > gpurun_out/c2_tma_overlap.json 2> gpurun_out/c2_tma_overlap.err; cat gpurun_out/c2_tma_overlap.json; tail -2 gpurun_out/c2_tma_overlap.err
timeout 500 python -m pytest tests/test_gpu_resnet_parity.py -q -s -m gpu > gpurun_out/c2_parity.log 2>&1; tail -40 gpurun_out/c2_parity.log
timeout 200 python bench/dgrad_probe.py > gpurun_out/c2_dgrad_probe.jsonl 2> gpurun_out/c2_dgrad_probe.err; cat gpurun_out/c2_dgrad_probe.jsonl; tail -3 gpurun_out/c2_dgrad_probe.err
for v in nn tn_wt; do
  timeout 170 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel -s 1 -c 1 -f -o gpurun_out/c2_ncu_dgrad_$v \
     python bench/dgrad_probe.py --one 0 $v > gpurun_out/c2_ncu_$v.log 2>&1; tail -2 gpurun_out/c2_ncu_$v.log
done
