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


def test_halo_candidates_are_gated_and_self_checked():
    """The halo-load kernels are never offered for CPU tensors, and one wrong answer disables them for the process."""
    x = torch.empty(256, 64, 56, 56, dtype=torch.bfloat16)
    w3 = torch.empty(64, 64, 3, 3, dtype=torch.bfloat16)
    conv.set_halo(True)
    try:
        assert conv._halo_caps(x, w3, 1) == {"fprop": False, "fprop2": False, "dgrad": False, "dgrad2": False}   # CPU tensor
        ref = torch.randn(4, 8, 6, 6)
        assert conv._halo_check("fprop_th", (1, 2, 3), ref + 1e-4, ref)
        assert conv.halo_state()["enabled"] and conv.halo_state()["checked"] == 1
        bad = ref.clone(); bad[0, 0, 0, 0] += 10.0
        assert not conv._halo_check("fprop_th", (1, 2, 3), bad, ref)
        st = conv.halo_state()
        assert not st["enabled"] and st["failed"] == ["fprop_th:1x2x3"]
        nan = ref.clone(); nan[1, 1, 1, 1] = float("nan")
        assert not conv._close(nan, ref)
    finally:
        conv.set_halo(conv._HALO)            # back to the process default (SHIPYARD_CONV_HALO, on unless set to 0)
    assert conv.halo_state() == {"enabled": conv._HALO, "checked": 0, "failed": []}


def test_experimental_variant_table():
    """Which additional halo variants take part in the race per ResNet-50 layer shape (pure shape logic)."""
    e = conv.experimental_impls
    assert e(256, 64, 56, 56, 64, 3, 1) == {"fprop": ["tha", "th264", "th264a"], "dgrad": ["tha"], "wgrad": ["th"]}
    assert e(256, 128, 28, 28, 128, 3, 1) == {"fprop": ["th2w"], "dgrad": ["th2w"], "wgrad": ["th"]}
    assert e(256, 256, 14, 14, 256, 3, 1) == {"fprop": [], "dgrad": [], "wgrad": ["th"]}
    assert e(256, 512, 7, 7, 512, 3, 1) == {"fprop": [], "dgrad": [], "wgrad": ["th"]}
    assert e(256, 128, 56, 56, 128, 3, 2) == {"fprop": [], "dgrad": [], "wgrad": []}          # stride 2: not a halo shape
    assert e(256, 64, 56, 56, 256, 1, 1) == {"fprop": [], "dgrad": [], "wgrad": []}           # 1x1
    assert e(256, 128, 56, 56, 128, 3, 1)["fprop"] == []                                     # 56x56 box does not fit the 23 KB slots
    assert set(conv._HALO_KW) >= {"th", "th2", "tha", "th264", "th264a", "th2w"}
    assert conv._EXP                                        # validated on hardware in round 2: in the race unless SHIPYARD_CONV_EXPERIMENTAL=0


def test_race_tie_break_prefers_native_inside_the_noise_band():
    t = {"dgrad_cudnn": 100.0, "dgrad_tc": 102.0, "dgrad_tc2": 110.0}
    assert conv._pick(t, "dgrad_") == "dgrad_tc"            # 2 % slower: inside the measured repeatability of the race
    t["dgrad_tc"] = 106.0
    assert conv._pick(t, "dgrad_") == "dgrad_cudnn"         # outside the band the library wins
    assert conv._pick({"wgrad_cudnn": 50.0}, "wgrad_") == "wgrad_cudnn"
    assert conv._pick({"fprop_cudnn": 60.0, "fprop_tc_stats": 40.0, "fprop_th": 39.0}, "fprop_") == "fprop_th"
