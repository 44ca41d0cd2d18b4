"""tcgen05 GEMM numerics vs a plain PyTorch fp32 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(128, 64, 64), (256, 256, 64), (128 * 5 + 17, 256, 256), (4096, 128, 512), (1000, 1000, 2048), (300, 64, 136),
          (8192, 512, 1024), (50176, 256, 64), (77, 2048, 512), (128, 72, 64)]


@pytest.mark.parametrize("m,n,k", SHAPES)
@pytest.mark.parametrize("block_n", [0, 64, 128, 256])
def test_gemm_matches_fp32_reference(m, n, k, block_n):
    from batch_shipyard_b200.ops import gemm
    torch.manual_seed(m + n + k)
    a = (torch.randn(m, k, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(n, k, device="cuda") * 0.5).to(torch.bfloat16)
    out = gemm.gemm_tn(a, b, block_n=block_n)
    ref = a.float() @ b.float().t()
    torch.testing.assert_close(out.float(), ref, atol=0.02 * (k ** 0.5) * 0.25 + 0.02, rtol=2e-2)


def test_gemm_bias_stats_and_strides():
    from batch_shipyard_b200.ops import gemm
    torch.manual_seed(0)
    m, n, k = 3000, 192, 320
    big = (torch.randn(m, k + 24, device="cuda")).to(torch.bfloat16)
    a = big[:, 8:8 + k]                                    # strided view (lda != K)
    b = (torch.randn(n, k, device="cuda") * 0.3).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda").to(torch.bfloat16)
    out = gemm.gemm_tn(a, b, bias=bias)
    ref = a.float() @ b.float().t() + bias.float()
    torch.testing.assert_close(out.float(), ref, atol=0.15, rtol=2e-2)
    stats = torch.zeros(2 * n, dtype=torch.float32, device="cuda")
    out = gemm.gemm_tn(a, b, stats=stats)
    of = out.float()
    torch.testing.assert_close(stats[:n], of.sum(0), atol=0.5, rtol=2e-3)
    torch.testing.assert_close(stats[n:], (of * of).sum(0), atol=2.0, rtol=2e-3)
    # persistent scheduling with fewer CTAs than tiles exercises the accumulator ping-pong and stat flushes
    stats2 = torch.zeros_like(stats)
    out2 = gemm.gemm_tn(a, b, stats=stats2, block_n=64, max_ctas=3)
    assert torch.equal(out2, out)
    torch.testing.assert_close(stats2, stats, atol=0.5, rtol=1e-3)


def test_conv1x1_and_linear_autograd():
    from batch_shipyard_b200.ops import gemm
    import torch.nn.functional as F
    torch.manual_seed(1)
    x = torch.randn(4, 64, 14, 14, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(128, 64, 1, 1, device="cuda") * 0.1).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y, stats = gemm.conv1x1_nhwc(x, w, want_stats=True)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    yr = F.conv2d(xr, wr)
    torch.testing.assert_close(y.float(), yr, atol=0.06, rtol=2e-2)
    torch.testing.assert_close(stats[:128], y.float().sum(dim=(0, 2, 3)), atol=0.5, rtol=5e-3)
    g = torch.randn_like(y)
    y.backward(g); yr.backward(g.float())
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=0.08, rtol=3e-2)
    torch.testing.assert_close(w.grad.float(), wr.grad, atol=0.5, rtol=3e-2)
    xl = torch.randn(33, 2048, device="cuda").to(torch.bfloat16).requires_grad_(True)
    wl = (torch.randn(1000, 2048, device="cuda") * 0.02).to(torch.bfloat16).requires_grad_(True)
    bl = torch.randn(1000, device="cuda").to(torch.bfloat16).requires_grad_(True)
    yl = gemm.linear(xl, wl, bl)
    ref = F.linear(xl.detach().float(), wl.detach().float(), bl.detach().float())
    torch.testing.assert_close(yl.float(), ref, atol=0.08, rtol=2e-2)
    yl.sum().backward()
    torch.testing.assert_close(bl.grad.float(), torch.full((1000,), 33.0, device="cuda"), atol=0.5, rtol=1e-2)
