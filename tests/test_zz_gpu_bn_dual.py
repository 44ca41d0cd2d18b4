"""Residual-gradient fusion (two-handle BatchNorm outputs, dual-gradient backward kernels) vs plain PyTorch fp32 references.

Kept in a file that sorts after the other GPU tests: the opt-in path (SHIPYARD_BN_DUAL, off by default) was added late in round 1
and only its first parametrisation ran on hardware before the GPU budget ended (it passed its numerics; see
profiles/conv_halo.md "Round-23 run") — a regression here must not stop `pytest -x` before the established tests have run.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c,hw,res", [(256, 14, True), (64, 28, False), (1024, 7, True)])
def test_fused_bn_two_gradient_backward(c, hw, res):
    """dual=True hands the output out twice; the two incoming gradients are summed inside the backward kernels."""
    from batch_shipyard_b200.ops import fused
    torch.manual_seed(1)
    n = 5
    x = (torch.randn(n, c, hw, hw, device="cuda") * 1.5 + 0.3).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    r = torch.randn(n, c, hw, hw, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if res else None
    gamma = (torch.rand(c, device="cuda") + 0.5).to(torch.bfloat16)
    beta = (torch.randn(c, device="cuda") * 0.1).to(torch.bfloat16)
    xs = [t.clone().requires_grad_(True) for t in (x, gamma, beta)] + ([r.clone().requires_grad_(True)] if res else [None])
    ya, yb = fused.fused_bn_act(xs[0], xs[1], xs[2], xs[3], None, None, True, dual=True)
    assert ya.data_ptr() == yb.data_ptr()
    g1 = torch.randn_like(ya); g2 = torch.randn_like(ya)
    torch.autograd.backward([ya, yb], [g1, g2])
    xr = [t.detach().float().requires_grad_(True) for t in (x, gamma, beta)] + ([r.detach().float().requires_grad_(True)] if res else [None])
    ref = fused.bn_act_reference(xr[0], xr[1], xr[2], xr[3], True)
    ref.backward(g1.float() + g2.float())
    torch.testing.assert_close(ya.float(), ref, atol=4e-2, rtol=2e-2)
    torch.testing.assert_close(xs[0].grad.float(), xr[0].grad, atol=8e-2, rtol=5e-2)
    m = n * hw * hw
    torch.testing.assert_close(xs[1].grad.float(), xr[1].grad, atol=0.03 * m ** 0.5 + 0.5, rtol=3e-2)
    torch.testing.assert_close(xs[2].grad.float(), xr[2].grad, atol=0.03 * m ** 0.5 + 0.5, rtol=3e-2)
    if res:
        torch.testing.assert_close(xs[3].grad.float(), xr[3].grad, atol=3e-2, rtol=2e-2)
    # only one handle used: the other gradient arrives as None and the single-gradient kernels run
    x2 = x.clone().requires_grad_(True)
    ya, yb = fused.fused_bn_act(x2, gamma, beta, r, None, None, True, dual=True)
    ya.backward(g1)
    x3 = x.clone().requires_grad_(True)
    fused.fused_bn_act(x3, gamma, beta, r, None, None, True).backward(g1)
    # same kernels, but the per-channel sums are accumulated with float atomics: equal up to summation order
    torch.testing.assert_close(x2.grad.float(), x3.grad.float(), atol=2e-2, rtol=2e-2)


def test_resnet_residual_gradient_fusion_matches_unfused():
    """Same model, same input: the gradients with the two-handle block outputs (SHIPYARD_BN_DUAL) are as close to an fp32 run of
    the model as the unfused bf16 gradients are (a tiny random network has a large bf16 noise floor, so the two bf16 runs are
    not compared with each other but each against fp32; the full ResNet-50 version is tests/test_gpu_resnet_parity.py)."""
    import copy
    from batch_shipyard_b200.models import resnet
    torch.manual_seed(3)
    model = resnet.ResNet((2, 2, 1, 1), 10, width=16).cuda().train()
    for m_ in model.modules():
        if isinstance(m_, resnet.ConvBN):
            m_.gamma.data.uniform_(0.5, 1.5)           # the zero-initialised last gamma of each block would hide the c1/c2 gradients
    for p_ in model.parameters():                      # parameters in bf16, BatchNorm running statistics stay fp32 (kernel contract)
        p_.data = p_.data.to(torch.bfloat16)
    ref = copy.deepcopy(model).float()                 # fp32 parameters and input -> the plain PyTorch path of ConvBN
    x = torch.randn(8, 3, 64, 64, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (8,), device="cuda")
    lr_ = torch.nn.functional.cross_entropy(ref(x.float()), y); lr_.backward()
    gref = torch.cat([p.grad.flatten() for p in ref.parameters()])
    grads = {}
    try:
        for dual in (False, True):
            resnet.set_bn_dual(dual)
            model.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(model(x).float(), y)
            loss.backward()
            grads[dual] = (float(loss), torch.cat([p.grad.float().flatten() for p in model.parameters()]))
    finally:
        resnet.set_bn_dual(False)
    assert abs(grads[True][0] - float(lr_)) < 0.05 * max(1.0, abs(float(lr_))) and abs(grads[False][0] - float(lr_)) < 0.05 * max(1.0, abs(float(lr_)))
    err = {d: float((grads[d][1] - gref).norm() / gref.norm()) for d in (False, True)}
    assert err[True] <= 1.5 * err[False] + 0.02, err


def test_maxpool_bwd2_variant_in_subprocess():
    """Both max-pool backward kernels (2x2-block = default, per-pixel = SHIPYARD_MAXPOOL_BWD2=0; read once per process by the library) must
    reproduce autograd's gradient."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, torch; sys.path.insert(0, %r)\n"
        "from batch_shipyard_b200.ops import fused\n"
        "import torch.nn.functional as F\n"
        "torch.manual_seed(0)\n"
        "for shape in [(3, 64, 24, 30), (8, 64, 112, 112), (2, 16, 6, 4)]:\n"
        "    x = torch.randn(*shape, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)\n"
        "    y = fused.maxpool3x3s2(x); g = torch.randn_like(y); y.backward(g)\n"
        "    xr = x.detach().float().requires_grad_(True); yr = F.max_pool2d(xr, 3, 2, 1); yr.backward(g.float())\n"
        "    assert torch.equal(y.float(), yr)\n"
        "    torch.testing.assert_close(x.grad.float(), xr.grad, atol=2e-2, rtol=2e-2)\n"
        "print('bwd2 ok')\n") % root
    for flag in ("1", "0"):
        p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SHIPYARD_MAXPOOL_BWD2=flag), stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=300)
        assert p.returncode == 0 and "bwd2 ok" in p.stdout, (flag, p.stdout[-3000:])
