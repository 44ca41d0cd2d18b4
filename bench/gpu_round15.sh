#!/bin/bash
# 1-GPU round: exactly what the driver does at round end (full GPU suite, smoke, default bench) + ncu capture of the tcgen05 kernels.
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/pytest_all15.log 2>&1; tail -6 gpurun_out/pytest_all15.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke15.log
timeout 300 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench15_default.log
cat > /tmp/gemm_one.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from batch_shipyard_b200.ops import gemm
a = torch.randn(8192, 8192, device='cuda').to(torch.bfloat16); b = torch.randn(8192, 8192, device='cuda').to(torch.bfloat16)
out = torch.empty(8192, 8192, device='cuda', dtype=torch.bfloat16)
x = torch.randn(256, 256, 28, 28, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = (torch.randn(256, 256, 3, 3, device='cuda') * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
dy = torch.randn(256, 256, 28, 28, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
st = torch.zeros(512, device='cuda')
for _ in range(2):
    gemm.gemm_tn(a, b, out=out, two_cta=True)
    gemm.gemm_tn(a, b, out=out)
    gemm.conv_fprop_nhwc(x, w, 1, 1, stats=st, two_cta=True)
    gemm.conv_dgrad_nhwc(dy, w, 1, two_cta=True)
    gemm.conv_wgrad_nhwc(x, dy, w.shape, 1, 1)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 5 -o gpurun_out/ncu_gemm15 -f python /tmp/gemm_one.py > gpurun_out/ncu_gemm15.log 2>&1
timeout 120 ncu -i gpurun_out/ncu_gemm15.ncu-rep --page raw --csv > gpurun_out/ncu_gemm15_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_gemm15.log
