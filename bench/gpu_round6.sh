#!/bin/bash
# 1-GPU round: new MN-major GEMM kernels, BN kernel bandwidth, HPCG, ncu captures for profiles/, bench.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "nn_mn or nt_wgrad or existing_grad or autograd" 2>&1 | tail -15 | tee gpurun_out/pytest_gemm6.log
timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_hpcg.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_ops6.log
timeout 200 python bench/wgrad_bench.py 2>&1 | tee gpurun_out/wgrad_bench6.log
timeout 200 python bench/bn_bench.py 2>&1 | tee gpurun_out/bn_bench6.log
timeout 120 python bench/bn_bench.py --blocks-per-sm 2 2>&1 | tail -1 | tee -a gpurun_out/bn_bench6.log
timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench6_n1.log
timeout 200 python recipes/HPCG-Infiniband-IntelMPI/run_hpcg.py --n 128 --t 5 2>&1 | tail -2 | tee gpurun_out/hpcg6_n1.log
# ncu: one capture of the top kernels (BN backward pair + forward) on the largest layer shape, full set, for profiles/
cat > /tmp/bn_one.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from batch_shipyard_b200.ops import fused
x = torch.randn(256, 56, 56, 256, device='cuda', dtype=torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
g = torch.nn.Parameter(torch.ones(256, device='cuda', dtype=torch.bfloat16)); b = torch.nn.Parameter(torch.zeros(256, device='cuda', dtype=torch.bfloat16))
for _ in range(2):
    y = fused.fused_bn_act(x, g, b, None, None, None, relu=True)
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bn_ -s 4 -c 4 -o gpurun_out/ncu_bn6 -f python /tmp/bn_one.py > gpurun_out/ncu_bn6.log 2>&1
timeout 120 ncu -i gpurun_out/ncu_bn6.ncu-rep --page raw --csv > gpurun_out/ncu_bn6_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_bn6.log
