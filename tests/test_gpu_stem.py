"""tcgen05 stem kernels (native/gemm/stem_s2d.inc) against a plain PyTorch fp32 reference of the same op: the dense 4x4 convolution
over the 16-channel space-to-depth input (forward + BatchNorm statistics) and its weight gradient."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

SHAPES = [(2, 19, 27), (4, 35, 35), (3, 115, 115), (1, 4, 128), (5, 23, 68)]       # (N, Hp, Wp)


def _data(n, hp, wp, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(n, 16, hp, wp, device="cuda", generator=g) * 0.7).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 16, 4, 4, device="cuda", generator=g) * 0.08).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    return x, w


@pytest.mark.parametrize("n,hp,wp", SHAPES)
@pytest.mark.parametrize("max_ctas", [0, 3])
def test_stem_fprop_and_stats(n, hp, wp, max_ctas):
    from batch_shipyard_b200.ops import gemm
    x, w = _data(n, hp, wp)
    ref = F.conv2d(x.float(), w.float())
    stats = torch.zeros(128, dtype=torch.float32, device="cuda")
    y = gemm.stem_s2d_fprop(x, w, stats=stats, max_ctas=max_ctas)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(y.float(), ref, atol=3e-2, rtol=2e-2)
    yb = y.float()                                           # the statistics are those of the bf16 output (what BatchNorm normalises)
    torch.testing.assert_close(stats[:64], yb.sum(dim=(0, 2, 3)), atol=0.5, rtol=2e-3)
    torch.testing.assert_close(stats[64:], (yb * yb).sum(dim=(0, 2, 3)), atol=0.5, rtol=2e-3)
    y2 = gemm.stem_s2d_fprop(x, w, max_ctas=max_ctas)          # without statistics: identical output
    assert torch.equal(y, y2)


@pytest.mark.parametrize("n,hp,wp", SHAPES)
@pytest.mark.parametrize("max_ctas", [0, 2])
def test_stem_wgrad(n, hp, wp, max_ctas):
    from batch_shipyard_b200.ops import gemm
    x, w = _data(n, hp, wp, seed=1)
    g = torch.Generator(device="cuda").manual_seed(2)
    dy = (torch.randn(n, 64, hp - 3, wp - 3, device="cuda", generator=g) * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wr = w.float().requires_grad_(True)
    F.conv2d(x.float(), wr).backward(dy.float())
    ref = wr.grad
    dw = gemm.stem_s2d_wgrad(x, dy, max_ctas=max_ctas)
    assert dw.shape == ref.shape
    scale = float(ref.abs().max())
    torch.testing.assert_close(dw.float(), ref, atol=2e-2 * scale, rtol=2e-2)
    # accumulate into an existing gradient; the workspace is left clean (a second call gives the same answer)
    out = dw.permute(0, 2, 3, 1).contiguous().clone()
    dw2 = gemm.stem_s2d_wgrad(x, dy, out=out, accumulate=True, max_ctas=max_ctas)
    torch.testing.assert_close(dw2.float(), 2 * ref, atol=4e-2 * scale, rtol=3e-2)
    ws, tickets = gemm._workspace(x.device)
    assert float(ws.abs().max()) == 0.0 and int(tickets.abs().max()) == 0


def test_stem_autograd_in_convbn_matches_cudnn_path():
    """ConvBN stem on the native kernels vs the same module on F.conv2d (cuDNN): outputs and weight gradient."""
    from batch_shipyard_b200.models import resnet
    from batch_shipyard_b200.ops import conv, fused
    torch.manual_seed(0)
    m = resnet.ConvBN(3, 64, 7, 2).cuda().train()
    for p in m.parameters():
        p.data = p.data.to(torch.bfloat16)
    img = torch.randint(0, 256, (8, 64, 96, 3), dtype=torch.uint8, device="cuda")
    s2d = torch.empty(8, 35, 51, 16, dtype=torch.bfloat16, device="cuda")
    fused.u8_to_s2d_norm(img, s2d)
    x = s2d.permute(0, 3, 1, 2)
    outs = {}
    for impl in ("cudnn", "tc"):
        conv.set_stem(impl)
        m.zero_grad(set_to_none=True)
        y = m(x)
        (y.float() ** 2).mean().backward()
        outs[impl] = (y.detach().float(), m.weight.grad.detach().float().clone(), m.gamma.grad.detach().float().clone())
    conv.set_stem("tc")
    torch.testing.assert_close(outs["tc"][0], outs["cudnn"][0], atol=6e-2, rtol=3e-2)
    for a, b in zip(outs["tc"][1:], outs["cudnn"][1:]):
        assert float((a - b).norm()) <= 0.03 * float(b.norm()) + 1e-5, (float((a - b).norm()), float(b.norm()))
