"""Fused-op numerics on the GPU vs plain PyTorch fp32 references."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_u8_to_bf16_norm():
    from batch_shipyard_b200.ops import fused
    x = torch.randint(0, 256, (5, 17, 19, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty(x.shape, dtype=torch.bfloat16, device="cuda")
    fused.u8_to_bf16_norm(x, out)
    mean = torch.tensor(fused.IMAGENET_MEAN, device="cuda"); std = torch.tensor(fused.IMAGENET_STD, device="cuda")
    ref = (x.float() / 255.0 - mean) / std
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=1e-2)


@pytest.mark.parametrize("c,hw,relu,res", [(64, 56, True, False), (256, 14, True, True), (2048, 7, False, False), (16, 9, True, True)])
def test_fused_bn_fwd_bwd(c, hw, relu, res):
    from batch_shipyard_b200.ops import fused
    torch.manual_seed(0)
    n = 6
    x = (torch.randn(n, c, hw, hw, device="cuda") * 1.5 + 0.3).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    r = torch.randn(n, c, hw, hw, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if res else None
    gamma = (torch.rand(c, device="cuda") + 0.5).to(torch.bfloat16)
    beta = (torch.randn(c, device="cuda") * 0.1).to(torch.bfloat16)
    rm = torch.zeros(c, device="cuda"); rv = torch.ones(c, device="cuda")
    xs = [t.clone().requires_grad_(True) for t in (x, gamma, beta)] + ([r.clone().requires_grad_(True)] if res else [None])
    out = fused.fused_bn_act(xs[0], xs[1], xs[2], xs[3], rm, rv, relu)
    dout = torch.randn_like(out)
    out.backward(dout)
    # fp32 reference of the same op
    xr = [t.detach().float().requires_grad_(True) for t in (x, gamma, beta)] + ([r.detach().float().requires_grad_(True)] if res else [None])
    ref = fused.bn_act_reference(xr[0], xr[1], xr[2], xr[3], relu)
    ref.backward(dout.float())
    torch.testing.assert_close(out.float(), ref, atol=4e-2, rtol=2e-2)
    torch.testing.assert_close(xs[0].grad.float(), xr[0].grad, atol=6e-2, rtol=5e-2)
    m = n * hw * hw
    torch.testing.assert_close(xs[1].grad.float(), xr[1].grad, atol=0.02 * m ** 0.5 + 0.5, rtol=3e-2)
    torch.testing.assert_close(xs[2].grad.float(), xr[2].grad, atol=0.02 * m ** 0.5 + 0.5, rtol=3e-2)
    if res:
        torch.testing.assert_close(xs[3].grad.float(), xr[3].grad, atol=2e-2, rtol=2e-2)
    # running statistics follow torch semantics (momentum 0.1, unbiased variance)
    xf = x.float()
    torch.testing.assert_close(rm, 0.1 * xf.mean(dim=(0, 2, 3)), atol=1e-2, rtol=1e-2)
    torch.testing.assert_close(rv, 0.9 + 0.1 * xf.var(dim=(0, 2, 3), unbiased=True), atol=2e-2, rtol=2e-2)


def test_maxpool_matches_torch():
    from batch_shipyard_b200.ops import fused
    import torch.nn.functional as F
    torch.manual_seed(0)
    x = torch.randn(3, 64, 23, 30, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = fused.maxpool3x3s2(x)
    xr = x.detach().float().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    assert torch.equal(y.float(), yr)
    g = torch.randn_like(y)
    y.backward(g); yr.backward(g.float())
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=2e-2, rtol=2e-2)


def test_s2d_stem_equivalence():
    from batch_shipyard_b200.ops import fused
    import torch.nn.functional as F
    torch.manual_seed(0)
    img = torch.randint(0, 256, (2, 32, 48, 3), dtype=torch.uint8, device="cuda")
    s2d = torch.empty(2, 19, 27, 16, dtype=torch.bfloat16, device="cuda")
    fused.u8_to_s2d_norm(img, s2d)
    mean = torch.tensor(fused.IMAGENET_MEAN, device="cuda"); std = torch.tensor(fused.IMAGENET_STD, device="cuda")
    xn = ((img.float() / 255.0 - mean) / std).permute(0, 3, 1, 2)
    torch.testing.assert_close(s2d.permute(0, 3, 1, 2).float(), fused.s2d_reference(xn), atol=2e-2, rtol=1e-2)
    w = torch.randn(8, 3, 7, 7, device="cuda") * 0.1
    ref = F.conv2d(xn, w, stride=2, padding=3)
    out = F.conv2d(s2d.permute(0, 3, 1, 2).float(), fused.stem_weight_s2d(w))
    torch.testing.assert_close(out, ref, atol=5e-2, rtol=2e-2)


def test_trainer_matches_reference_sgd():
    """Two steps of the fused trainer (tiny ResNet) track a plain fp32 PyTorch SGD run."""
    import copy
    import torch.nn.functional as F
    from batch_shipyard_b200.models.resnet import resnet_tiny
    from batch_shipyard_b200.ops.coll import Communicator
    from batch_shipyard_b200.parallel.ddp import FusedDataParallelTrainer
    torch.manual_seed(0)
    model = resnet_tiny(10)
    ref = copy.deepcopy(model).cuda().float()
    comm = Communicator(0, 1, device=0, heap_bytes=256 << 20)
    tr = FusedDataParallelTrainer(model, comm, (16, 3, 64, 64), 10, lr=0.05, momentum=0.9, weight_decay=1e-4, use_graph=False)
    x = torch.randn(16, 3, 64, 64, device="cuda")
    y = torch.randint(0, 10, (16,), device="cuda")
    img = torch.randint(0, 256, (16, 64, 64, 3), dtype=torch.uint8, device="cuda")
    tr.load_images_u8(img, y)
    opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    ref.train()
    from batch_shipyard_b200.ops import fused
    mean = torch.tensor(fused.IMAGENET_MEAN, device="cuda"); std = torch.tensor(fused.IMAGENET_STD, device="cuda")
    xr = ((img.float() / 255.0 - mean) / std).permute(0, 3, 1, 2).contiguous()
    for _ in range(2):
        loss = float(tr.step())
        opt.zero_grad(); lr = F.cross_entropy(ref(xr), y); lr.backward(); opt.step()
        assert abs(loss - float(lr)) < 0.15 * max(1.0, abs(float(lr))), (loss, float(lr))
    comm.check_status()
    assert float(tr.flat.grads.abs().max()) == 0.0
    comm.close()
