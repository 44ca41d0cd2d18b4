"""Halo-load 3x3 convolution kernels (native/gemm/conv_halo.inc) vs a plain PyTorch fp32 reference of the same op."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [  # n, channels in, h, w, channels out
    (4, 64, 56, 56, 64),        # weights-stationary path (one column block, one channel block), R = 2
    (8, 128, 28, 28, 128),      # R = 4, two channel blocks
    (8, 256, 14, 14, 256),      # R = 7, 16-pixel padded rows
    (16, 512, 7, 7, 512),       # R = 7, two column blocks at BN = 256
    (2, 64, 28, 28, 192),       # partial last column block (BN = 256 > Cout)
    (6, 128, 12, 20, 64),       # non-square image, Cin != Cout
]


def _mk(n, cin, h, w, cout):
    torch.manual_seed(n + cin + h + cout)
    x = (torch.randn(n, cin, h, w, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, 3, 3, device="cuda") * (1.0 / (cin * 9) ** 0.5)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    return x, wt


@pytest.mark.parametrize("n,cin,h,w,cout", CASES)
@pytest.mark.parametrize("pair", [False, True])
def test_halo_fprop_stats_and_dgrad(n, cin, h, w, cout, pair):
    from batch_shipyard_b200.ops import gemm
    import os
    if pair and os.environ.get("SHIPYARD_HALO_PAIR", "1") == "0":
        pytest.skip("CTA-pair halo kernels disabled (SHIPYARD_HALO_PAIR=0)")
    if pair and not gemm.halo_ok(n, h, w, cin, cout, 3, 3, 1, 1, pair=True):
        pytest.skip("CTA pairs need out channels % 128 and an even number of M tiles")
    x, wt = _mk(n, cin, h, w, cout)
    ref = F.conv2d(x.float(), wt.float(), padding=1)
    y = gemm.conv3x3_halo(x, wt, pair=pair)
    torch.testing.assert_close(y.float(), ref, atol=0.03, rtol=2e-2)
    stats = torch.zeros(2 * cout, dtype=torch.float32, device="cuda")
    y2 = gemm.conv3x3_halo(x, wt, stats=stats, pair=pair)
    assert torch.equal(y2, y)
    torch.testing.assert_close(stats[:cout], y.float().sum(dim=(0, 2, 3)), atol=0.5, rtol=5e-3)
    torch.testing.assert_close(stats[cout:], (y.float() ** 2).sum(dim=(0, 2, 3)), atol=0.5, rtol=5e-3)
    if cin % 64 == 0 and (not pair or gemm.halo_ok(n, h, w, cout, cin, 3, 3, 1, 1, pair=True, dgrad=True)) and cout % 64 == 0:
        dy = (torch.randn_like(ref) * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dx = gemm.conv3x3_halo(dy, wt, dgrad=True, pair=pair)
        dref = torch.nn.grad.conv2d_input(x.shape, wt.float(), dy.float(), stride=1, padding=1)
        torch.testing.assert_close(dx.float(), dref, atol=0.05, rtol=2e-2)


def test_halo_in_dispatcher_matches_library():
    """With the halo candidates enabled the dispatcher's forward/backward still reproduces the fp32 reference (whatever it picks),
    and every halo candidate it considered passed its self-check."""
    from batch_shipyard_b200.ops import conv
    conv.set_mode("auto")
    conv.set_halo(True)
    try:
        x, wt = _mk(32, 128, 28, 28, 128)
        x.requires_grad_(True); wt.requires_grad_(True)
        y, _ = conv.conv_bn_input(x, wt, 1)
        g = (torch.randn_like(y) * 0.5)
        y.backward(g)
        xr, wr = x.detach().float().requires_grad_(True), wt.detach().float().requires_grad_(True)
        F.conv2d(xr, wr, padding=1).backward(g.float())
        torch.testing.assert_close(x.grad.float(), xr.grad, atol=0.05, rtol=2e-2)
        st = conv.halo_state()
        assert st["checked"] >= 2 and st["failed"] == [], st
        conv.set_mode("tc")
        conv.set_halo(True)
        x2, w2 = _mk(32, 128, 28, 28, 128)
        plan = conv.plan_for(x2, w2, 1)
        import os
        want = "th" if os.environ.get("SHIPYARD_HALO_PAIR", "1") == "0" else "th2"
        assert plan.fprop == want and plan.dgrad == want
        y2, s2 = conv.conv_bn_input(x2, w2, 1)
        torch.testing.assert_close(y2.float(), F.conv2d(x2.float(), w2.float(), padding=1), atol=0.03, rtol=2e-2)
        assert s2 is not None
    finally:
        conv.set_halo(conv._HALO)            # back to the process default
        conv.set_mode("auto")


@pytest.mark.parametrize("variant", ["epi_alt", "weights_stationary", "pair64", "pair64_alt"])
def test_halo_unverified_variants(variant):
    """Alternate-tile epilogue (BN = 64) and all-weights-stationary CTA pairs (128 channels, 28x28): same numerics as the base kernels."""
    from batch_shipyard_b200.ops import gemm
    if variant == "epi_alt":
        n, cin, h, w, cout, kw = 8, 64, 56, 56, 64, dict(epi_alt=True)
    elif variant.startswith("pair64"):                   # 64-column CTA pairs (fprop only)
        n, cin, h, w, cout, kw = 8, 64, 56, 56, 64, dict(pair=True, block_n=64, epi_alt=variant.endswith("alt"))
    else:
        n, cin, h, w, cout, kw = 8, 128, 28, 28, 128, dict(pair=True, weights_stationary=True)
    x, wt = _mk(n, cin, h, w, cout)
    ref = F.conv2d(x.float(), wt.float(), padding=1)
    stats = torch.zeros(2 * cout, dtype=torch.float32, device="cuda")
    y = gemm.conv3x3_halo(x, wt, stats=stats, **kw)
    torch.testing.assert_close(y.float(), ref, atol=0.03, rtol=2e-2)
    torch.testing.assert_close(stats[:cout], y.float().sum(dim=(0, 2, 3)), atol=0.5, rtol=5e-3)
    torch.testing.assert_close(stats[cout:], (y.float() ** 2).sum(dim=(0, 2, 3)), atol=0.5, rtol=5e-3)
    assert torch.equal(gemm.conv3x3_halo(x, wt, **kw), y)
    if variant.startswith("pair64"):
        return
    dy = (torch.randn_like(ref) * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dx = gemm.conv3x3_halo(dy, wt, dgrad=True, **kw)
    dref = torch.nn.grad.conv2d_input(x.shape, wt.float(), dy.float(), stride=1, padding=1)
    torch.testing.assert_close(dx.float(), dref, atol=0.05, rtol=2e-2)


@pytest.mark.parametrize("n,cin,h,w,cout", [(4, 64, 56, 56, 64), (8, 128, 28, 28, 128), (8, 256, 14, 14, 128), (16, 128, 7, 7, 256), (6, 64, 12, 20, 64)])
@pytest.mark.parametrize("splits", [0, 1, 3])
def test_halo_wgrad_unverified(n, cin, h, w, cout, splits):
    from batch_shipyard_b200.ops import gemm
    x, wt = _mk(n, cin, h, w, cout)
    dy = (torch.randn(n, cout, h, w, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ref = torch.nn.grad.conv2d_weight(x.float(), wt.shape, dy.float(), padding=1)
    tol = 0.02 * (n * h * w) ** 0.5 * 0.25 + 0.05
    dw = gemm.conv3x3_wgrad_halo(x, dy, splits=splits)
    torch.testing.assert_close(dw.float(), ref, atol=tol, rtol=3e-2)
    base = torch.randn(cout, 3, 3, cin, device="cuda").to(torch.bfloat16)
    buf = base.clone()
    gemm.conv3x3_wgrad_halo(x, dy, out=buf, accumulate=True, splits=splits)
    torch.testing.assert_close(buf.permute(0, 3, 1, 2).float(), ref + base.permute(0, 3, 1, 2).float(), atol=tol + 0.05, rtol=3e-2)
    ws, tickets = gemm._workspace(x.device)
    assert float(ws.abs().max()) == 0.0 and int(tickets.abs().max()) == 0
