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


NN_SHAPES = [(256, 64, 64), (1000, 256, 64), (4096, 64, 256), (3000, 512, 128), (777, 128, 1000), (50176, 64, 256), (640, 2048, 512),
             (256, 2048, 1000), (512, 192, 320)]


@pytest.mark.parametrize("m,n,k", NN_SHAPES)
def test_gemm_nn_mn_major_b(m, n, k):
    """dgrad shape: B given as [K, N] row-major and consumed as an MN-major tcgen05 operand."""
    from batch_shipyard_b200.ops import gemm
    torch.manual_seed(m + 3 * n + k)
    a = (torch.randn(m, k, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(k, n, device="cuda") * 0.5).to(torch.bfloat16)
    out = gemm.gemm_nn(a, b)
    ref = a.float() @ b.float()
    torch.testing.assert_close(out.float(), ref, atol=0.02 * (k ** 0.5) * 0.25 + 0.02, rtol=2e-2)


NT_SHAPES = [(256, 64, 64), (4096, 64, 256), (4096, 256, 64), (50176, 128, 512), (12544, 512, 2048), (12544, 2048, 512), (256, 2048, 1000),
             (1000, 72, 200), (200704, 64, 64), (333, 128, 128)]


@pytest.mark.parametrize("k,i,j", NT_SHAPES)
@pytest.mark.parametrize("splits", [0, 1, 3])
def test_gemm_nt_wgrad_splitk(k, i, j, splits):
    """wgrad shape: out[J, I] = B[K,J]^T A[K,I]; split-K with in-kernel finalisation; workspace must come back zeroed."""
    from batch_shipyard_b200.ops import gemm
    torch.manual_seed(k + i + j)
    a = (torch.randn(k, i, device="cuda") * 0.25).to(torch.bfloat16)
    b = (torch.randn(k, j, device="cuda") * 0.25).to(torch.bfloat16)
    ref = b.float().t() @ a.float()
    out = gemm.gemm_nt_wgrad(a, b, splits=splits)
    tol = 0.01 * (k ** 0.5) * 0.0625 * 4 + 0.02
    torch.testing.assert_close(out.float(), ref, atol=tol, rtol=2e-2)
    ws, tickets = gemm._workspace(a.device)
    assert float(ws.abs().max()) == 0.0 and int(tickets.abs().max()) == 0
    # accumulate into an existing gradient buffer (run twice: workspace reuse across launches)
    base = torch.randn(j, i, device="cuda").to(torch.bfloat16)
    out2 = base.clone()
    gemm.gemm_nt_wgrad(a, b, out=out2, accumulate=True, splits=splits)
    torch.testing.assert_close(out2.float(), ref + base.float(), atol=tol + 0.05, rtol=3e-2)


def test_conv1x1_wgrad_written_into_existing_grad():
    """With a pre-existing .grad (the flat gradient buffer of the trainer) the kernels write into it directly."""
    from batch_shipyard_b200.ops import gemm
    import torch.nn.functional as F
    torch.manual_seed(2)
    x = torch.randn(8, 256, 14, 14, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    flat = torch.zeros(64 * 256, device="cuda", dtype=torch.bfloat16)
    w = torch.nn.Parameter((torch.randn(64, 256, 1, 1, device="cuda") * 0.1).to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    w.grad = flat.view(64, 1, 1, 256).permute(0, 3, 1, 2)
    y, _ = gemm.conv1x1_nhwc(x, w, want_stats=False)
    g = torch.randn_like(y)
    y.backward(g)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    F.conv2d(xr, wr).backward(g.float())
    torch.testing.assert_close(flat.view(64, 256).float(), wr.grad.view(64, 256), atol=0.5, rtol=3e-2)
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=0.08, rtol=3e-2)
    assert w.grad.data_ptr() == flat.data_ptr()


CONV_CASES = [  # n, cin, h, w, cout, k, stride
    (8, 64, 16, 16, 64, 3, 1), (8, 128, 28, 28, 128, 3, 1), (2, 256, 16, 16, 256, 3, 1), (16, 64, 8, 8, 192, 3, 1),
    (8, 128, 16, 16, 128, 3, 2), (8, 256, 16, 16, 512, 1, 2), (2, 64, 56, 56, 64, 3, 1), (32, 512, 4, 4, 512, 3, 1), (32, 64, 16, 16, 64, 3, 2)]


@pytest.mark.parametrize("n,cin,h,w,cout,k,stride", CONV_CASES)
def test_conv_implicit_gemm_fprop_and_dgrad(n, cin, h, w, cout, k, stride):
    """TMA-im2col implicit GEMM vs F.conv2d in fp32 (padding handled by TMA zero fill, taps by im2col offsets)."""
    from batch_shipyard_b200.ops import gemm
    import torch.nn.functional as F
    torch.manual_seed(n + cin + h + cout + k)
    pad = k // 2
    x = (torch.randn(n, cin, h, w, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, device="cuda") * (1.0 / (cin * k * k) ** 0.5)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert gemm.conv_supported(x, wt, stride, pad)
    stats = torch.zeros(2 * cout, dtype=torch.float32, device="cuda")
    y = gemm.conv_fprop_nhwc(x, wt, stride, pad, stats=stats)
    ref = F.conv2d(x.float(), wt.float(), stride=stride, padding=pad)
    torch.testing.assert_close(y.float(), ref, atol=0.03, rtol=2e-2)
    torch.testing.assert_close(stats[:cout], y.float().sum(dim=(0, 2, 3)), atol=0.5, rtol=5e-3)
    torch.testing.assert_close(stats[cout:], (y.float() ** 2).sum(dim=(0, 2, 3)), atol=0.5, rtol=5e-3)
    dy = (torch.randn_like(ref) * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    if stride == 1 and cout % 64 == 0:
        dx = gemm.conv_dgrad_nhwc(dy, wt, pad)
        dref = torch.nn.grad.conv2d_input(x.shape, wt.float(), dy.float(), stride=1, padding=pad)
        torch.testing.assert_close(dx.float(), dref, atol=0.05, rtol=2e-2)
    # weight gradient: all taps in one split-K launch, then again accumulating into an existing buffer
    wref = torch.nn.grad.conv2d_weight(x.float(), wt.shape, dy.float(), stride=stride, padding=pad)
    tol = 0.02 * (n * ref.shape[2] * ref.shape[3]) ** 0.5 * 0.25 + 0.05
    dw = gemm.conv_wgrad_nhwc(x, dy, wt.shape, stride, pad)
    torch.testing.assert_close(dw.float(), wref, atol=tol, rtol=3e-2)
    base = torch.randn(cout, k, k, cin, device="cuda").to(torch.bfloat16)
    buf = base.clone()
    gemm.conv_wgrad_nhwc(x, dy, wt.shape, stride, pad, out=buf, accumulate=True, splits=3)
    torch.testing.assert_close(buf.permute(0, 3, 1, 2).float(), wref + base.permute(0, 3, 1, 2).float(), atol=tol + 0.05, rtol=3e-2)
    ws, tickets = gemm._workspace(x.device)
    assert float(ws.abs().max()) == 0.0 and int(tickets.abs().max()) == 0


def test_conv_autograd_matches_cudnn():
    from batch_shipyard_b200.ops import gemm
    import torch.nn.functional as F
    torch.manual_seed(5)
    x = torch.randn(8, 64, 16, 16, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(128, 64, 3, 3, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y, _ = gemm.conv_nhwc(x, w, 1, 1, want_stats=True)
    g = torch.randn_like(y)
    y.backward(g)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    F.conv2d(xr, wr, padding=1).backward(g.float())
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=0.1, rtol=3e-2)
    torch.testing.assert_close(w.grad.float(), wr.grad, atol=0.6, rtol=3e-2)


@pytest.mark.parametrize("mode", ["tc", "cudnn", "auto"])
def test_conv_dispatcher_modes_agree_with_reference(mode):
    """ops.conv picks per pass between the tcgen05 kernels and cuDNN; every mode must give the same gradients."""
    from batch_shipyard_b200.ops import conv
    import torch.nn.functional as F
    conv.set_mode(mode)
    try:
        torch.manual_seed(11)
        for (n, cin, hw, cout, k, stride) in [(8, 64, 16, 64, 3, 1), (8, 256, 16, 64, 1, 1), (8, 128, 16, 256, 1, 2), (8, 128, 16, 128, 3, 2)]:
            x = torch.randn(n, cin, hw, hw, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            w = torch.nn.Parameter((torch.randn(cout, cin, k, k, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
            y, stats = conv.conv_bn_input(x, w, stride)
            g = torch.randn_like(y)
            y.backward(g)
            xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
            yr = F.conv2d(xr, wr, stride=stride, padding=k // 2)
            yr.backward(g.float())
            torch.testing.assert_close(y.float(), yr, atol=0.06, rtol=2e-2)
            torch.testing.assert_close(x.grad.float(), xr.grad, atol=0.1, rtol=3e-2)
            torch.testing.assert_close(w.grad.float(), wr.grad, atol=0.8, rtol=3e-2)
            if stats is not None:
                torch.testing.assert_close(stats[:cout], y.float().sum(dim=(0, 2, 3)), atol=0.5, rtol=5e-3)
        tab = conv.plan_table()
        assert len(tab) == 4
        if mode == "tc":
            assert all(v["fprop"] in ("tc", "tc2") and v["wgrad"] == "tc" for v in tab.values())
        if mode == "cudnn":
            assert all(v["fprop"] == "cudnn" and v["dgrad"] == "cudnn" and v["wgrad"] == "cudnn" for v in tab.values())
    finally:
        conv.set_mode("auto")


@pytest.mark.parametrize("m,n,k", [(256, 128, 64), (512, 256, 256), (4096, 512, 1024), (128 * 7 + 40, 256, 192), (8192, 1024, 512), (50176, 256, 64)])
@pytest.mark.parametrize("block_n", [0, 128])
def test_gemm_two_cta_pairs(m, n, k, block_n):
    """cta_group::2: one M=256 MMA per CTA pair, B split across the pair, multicast commits."""
    from batch_shipyard_b200.ops import gemm
    torch.manual_seed(m + n + k + 1)
    a = (torch.randn(m, k, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(n, k, device="cuda") * 0.5).to(torch.bfloat16)
    stats = torch.zeros(2 * n, dtype=torch.float32, device="cuda")
    out = gemm.gemm_tn(a, b, block_n=block_n, two_cta=True, stats=stats)
    ref = a.float() @ b.float().t()
    torch.testing.assert_close(out.float(), ref, atol=0.02 * (k ** 0.5) * 0.25 + 0.02, rtol=2e-2)
    torch.testing.assert_close(stats[:n], out.float().sum(0), atol=1.0, rtol=5e-3)
    out2 = gemm.gemm_tn(a, b, block_n=block_n, two_cta=True, max_ctas=4)      # persistent pairs: accumulator ping-pong across tiles
    assert torch.equal(out2, out)


@pytest.mark.parametrize("n,cin,h,w,cout,k,stride", [(8, 64, 16, 16, 128, 3, 1), (4, 128, 16, 16, 256, 3, 1), (16, 256, 8, 8, 256, 3, 1), (16, 128, 16, 16, 128, 3, 2),
                                                       (16, 256, 16, 16, 512, 1, 2)])
def test_conv_fprop_two_cta_pairs(n, cin, h, w, cout, k, stride):
    from batch_shipyard_b200.ops import gemm
    import torch.nn.functional as F
    torch.manual_seed(n + cin + cout)
    pad = k // 2
    x = (torch.randn(n, cin, h, w, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, device="cuda") * (1.0 / (cin * k * k) ** 0.5)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    stats = torch.zeros(2 * cout, dtype=torch.float32, device="cuda")
    y = gemm.conv_fprop_nhwc(x, wt, stride, pad, stats=stats, two_cta=True)
    ref = F.conv2d(x.float(), wt.float(), stride=stride, padding=pad)
    torch.testing.assert_close(y.float(), ref, atol=0.03, rtol=2e-2)
    torch.testing.assert_close(stats[:cout], y.float().sum(dim=(0, 2, 3)), atol=0.5, rtol=5e-3)


@pytest.mark.parametrize("m,n,k", [(512, 256, 64), (4096, 128, 512), (1024, 512, 2048), (50176, 256, 64)])
def test_gemm_nn_two_cta(m, n, k):
    from batch_shipyard_b200.ops import gemm
    torch.manual_seed(m + n + k + 2)
    a = (torch.randn(m, k, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(k, n, device="cuda") * 0.5).to(torch.bfloat16)
    out = gemm.gemm_nn(a, b, two_cta=True)
    torch.testing.assert_close(out.float(), a.float() @ b.float(), atol=0.02 * (k ** 0.5) * 0.25 + 0.02, rtol=2e-2)


@pytest.mark.parametrize("n,c,h,w,k", [(8, 128, 16, 16, 3), (4, 256, 16, 16, 3), (16, 128, 8, 8, 3)])
def test_conv_dgrad_two_cta(n, c, h, w, k):
    from batch_shipyard_b200.ops import gemm
    torch.manual_seed(n + c + h)
    cout = 128
    wt = (torch.randn(cout, c, k, k, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = (torch.randn(n, cout, h, w, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dx = gemm.conv_dgrad_nhwc(dy, wt, k // 2, two_cta=True)
    ref = torch.nn.grad.conv2d_input((n, c, h, w), wt.float(), dy.float(), stride=1, padding=k // 2)
    torch.testing.assert_close(dx.float(), ref, atol=0.05, rtol=2e-2)


def test_direct_store_epilogue_variants_in_subprocess():
    """SHIPYARD_GEMM_DIRECT_STORE=1 switches the TN / CTA-pair / im2col kernels to the st.global epilogue; the switch is read
    once per process, so the numerics tests of this file are re-run in a child process with it set."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ({}, {"SHIPYARD_GEMM_EPI_ALT": "1"}):              # second pass: alternate-tile epilogue on the 64-column GEMM tiles
        env = dict(os.environ, SHIPYARD_GEMM_DIRECT_STORE="1", **extra)
        p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_gemm.py"), "-x", "-q", "-m", "gpu", "-k",
                            "matches_fp32 or bias_stats or two_cta or conv_implicit or nn_mn_major or conv1x1_and_linear"],
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        assert p.returncode == 0, (extra, p.stdout[-4000:])


@pytest.mark.parametrize("n,cin,cout,p,q", [(4, 64, 128, 14, 14), (2, 256, 512, 28, 28), (3, 72, 40, 5, 9), (8, 1024, 2048, 7, 7)])
def test_conv1x1_stride2_dgrad_scatter_epilogue(n, cin, cout, p, q):
    """dX of a 1x1 / stride-2 convolution: GEMM + scattering st.global epilogue (row -> pixel (2p, 2q), zeros for the skipped pixels)
    against autograd in fp32; the output buffer is pre-filled with NaN-free garbage to prove every element is written exactly once."""
    from batch_shipyard_b200.ops import gemm
    torch.manual_seed(n + cin)
    dy = (torch.randn(n, cout, p, q, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(cout, cin, 1, 1, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x = torch.zeros(n, cin, 2 * p, 2 * q, device="cuda", requires_grad=True)
    torch.nn.functional.conv2d(x, w.float(), stride=2).backward(dy.float())
    for max_ctas in (0, 3):
        torch.full((n * 4 * p * q * cin,), 7.0, device="cuda", dtype=torch.bfloat16)       # dirty the allocator's next block
        dx = gemm.conv1x1_s2_dgrad(dy, w, max_ctas=max_ctas)
        assert dx.shape == x.shape
        torch.testing.assert_close(dx.float(), x.grad, atol=3e-2, rtol=2e-2)
        assert float(dx[:, :, 1::2, :].abs().max()) == 0.0 and float(dx[:, :, :, 1::2].abs().max()) == 0.0
