"""Convolution dispatcher on a CPU box: falls back to the library path, same numerics, no plan recorded."""
import torch
import torch.nn.functional as F

from batch_shipyard_b200.ops import conv


def test_cpu_falls_back_to_library_conv():
    conv.set_mode("auto")
    x = torch.randn(2, 8, 6, 6, requires_grad=True)
    w = torch.randn(4, 8, 3, 3, requires_grad=True)
    y, stats = conv.conv_bn_input(x, w, 1)
    assert stats is None
    ref = F.conv2d(x, w, padding=1)
    torch.testing.assert_close(y, ref)
    g = torch.randn_like(y)
    gx, gw = torch.autograd.grad(y, (x, w), g)
    rx, rw = torch.autograd.grad(ref, (x, w), g)
    torch.testing.assert_close(gx, rx)
    torch.testing.assert_close(gw, rw)
    assert conv.plan_table() == {}


def test_capabilities_table():
    x = torch.empty(256, 64, 56, 56, dtype=torch.bfloat16)
    w3 = torch.empty(64, 64, 3, 3, dtype=torch.bfloat16)
    caps = conv._tc_caps(x, w3, 1)
    assert caps == {"fprop": False, "dgrad": False, "wgrad": False}      # CPU tensors: never on the tcgen05 kernels
    plan = conv.ConvPlan()
    assert (plan.fprop, plan.dgrad, plan.wgrad, plan.stats) == ("cudnn", "cudnn", "cudnn", False)
