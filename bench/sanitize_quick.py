"""Tiny kernel selection for a time-boxed compute-sanitizer pass (bench/sanitize_gpu.sh is the full version)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from batch_shipyard_b200.ops import fused, gemm

torch.manual_seed(0)
x = torch.randn(4, 64, 8, 8, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
g = torch.nn.Parameter(torch.ones(64, device="cuda", dtype=torch.bfloat16)); b = torch.nn.Parameter(torch.zeros(64, device="cuda", dtype=torch.bfloat16))
y = fused.fused_bn_act(x, g, b, None, None, None, relu=True)
y.backward(torch.randn_like(y))
p = fused.maxpool3x3s2(y.detach().requires_grad_(True))
a = (torch.randn(256, 64, device="cuda") * 0.5).to(torch.bfloat16); w = (torch.randn(128, 64, device="cuda") * 0.5).to(torch.bfloat16)
st = torch.zeros(256, device="cuda")
o1 = gemm.gemm_tn(a, w, stats=st)
o2 = gemm.gemm_tn(a, w, two_cta=True)
o3 = gemm.gemm_nn(a, w.t().contiguous())
o4 = gemm.gemm_nt_wgrad(a, (torch.randn(256, 128, device="cuda") * 0.5).to(torch.bfloat16))
xc = (torch.randn(8, 64, 16, 16, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
wc = (torch.randn(64, 64, 3, 3, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
yc = gemm.conv_fprop_nhwc(xc, wc, 1, 1)
dx = gemm.conv_dgrad_nhwc(yc, wc, 1)
dw = gemm.conv_wgrad_nhwc(xc, yc, wc.shape, 1, 1)
torch.cuda.synchronize()
ref = a.float() @ w.float().t()
assert torch.allclose(o1.float(), ref, atol=0.3, rtol=2e-2) and torch.equal(o1, o2)
print("sanitize_quick: kernels ran, results consistent")
