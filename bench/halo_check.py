"""One-shot hardware check of the halo-load 3x3 convolution (native/gemm/conv_halo.inc).

`python bench/halo_check.py numerics <base_mode>`  — max relative error of every variant against F.conv2d (fp32), one JSON line each;
`python bench/halo_check.py timing <base_mode>`    — device time of halo / im2col / cuDNN on the four ResNet-50 3x3 shapes (batch 256).
base_mode 1 sets the A descriptor's base_offset field to (start_address >> 7) & 7 for the shifted taps, 0 leaves it zero: which of
the two the tensor core expects for a start address that is not 1024-byte aligned is the one thing the guides do not say.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from batch_shipyard_b200.ops import gemm

CASES = [(4, 64, 56, 56, 64), (8, 128, 28, 28, 128), (8, 256, 14, 14, 256), (16, 512, 7, 7, 512), (2, 64, 28, 28, 192), (6, 128, 12, 20, 64)]


def mk(n, cin, h, w, cout):
    torch.manual_seed(n + cin + h + cout)
    x = (torch.randn(n, cin, h, w, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, 3, 3, device="cuda") * (1.0 / (cin * 9) ** 0.5)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    return x, wt


def rel(a, ref):
    a, ref = a.float(), ref.float()
    if not bool(torch.isfinite(a).all()):
        return float("inf")
    return float((a - ref).abs().max()) / (float(ref.abs().max()) + 1e-9)


def numerics(mode: int) -> None:
    for pair in (False, True):                       # 1-CTA first: a trap in the pair kernels must not hide the 1-CTA answer
        for (n, cin, h, w, cout) in CASES:
            if pair and not gemm.halo_ok(n, h, w, cin, cout, 3, 3, 1, 1, pair=True):
                continue
            x, wt = mk(n, cin, h, w, cout)
            ref = F.conv2d(x.float(), wt.float(), padding=1)
            out = {"case": [n, cin, h, w, cout], "pair": pair, "base_mode": mode}
            try:
                y = gemm.conv3x3_halo(x, wt, pair=pair, base_mode=mode); torch.cuda.synchronize()
                out["fprop_rel"] = round(rel(y, ref), 5)
                st = torch.zeros(2 * cout, dtype=torch.float32, device="cuda")
                y2 = gemm.conv3x3_halo(x, wt, stats=st, pair=pair, base_mode=mode); torch.cuda.synchronize()
                out["stats_rel"] = round(max(rel(st[:cout], y2.float().sum(dim=(0, 2, 3))), rel(st[cout:], (y2.float() ** 2).sum(dim=(0, 2, 3)))), 5)
                if cout % 64 == 0 and (not pair or gemm.halo_ok(n, h, w, cout, cin, 3, 3, 1, 1, pair=True, dgrad=True)):
                    dy = (torch.randn_like(ref) * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
                    dx = gemm.conv3x3_halo(dy, wt, dgrad=True, pair=pair, base_mode=mode); torch.cuda.synchronize()
                    out["dgrad_rel"] = round(rel(dx, torch.nn.grad.conv2d_input(x.shape, wt.float(), dy.float(), stride=1, padding=1)), 5)
            except Exception as e:  # noqa: BLE001
                out["error"] = str(e)[:200]
                print(json.dumps(out), flush=True)
                return                                   # a CUDA error is sticky: nothing after it in this process means anything
            print(json.dumps(out), flush=True)


def t_us(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def timing(mode: int) -> None:
    for (cin, hw) in ((64, 56), (128, 28), (256, 14), (512, 7)):
        n, cout = 256, cin
        x, wt = mk(n, cin, hw, hw, cout)
        dy = torch.randn_like(x)
        st = torch.zeros(2 * cout, dtype=torch.float32, device="cuda")
        flops = 2.0 * n * hw * hw * cout * cin * 9
        row = {"shape": [n, cin, hw, hw, cout], "base_mode": mode}
        cands = {
            "fprop_cudnn": lambda: F.conv2d(x, wt, padding=1),
            "fprop_tc": lambda: gemm.conv_fprop_nhwc(x, wt, 1, 1, stats=st),
            "fprop_th": lambda: gemm.conv3x3_halo(x, wt, stats=st, base_mode=mode),
            "dgrad_cudnn": lambda: torch.ops.aten.convolution_backward(dy, x, wt, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False]),
            "dgrad_tc": lambda: gemm.conv_dgrad_nhwc(dy, wt, 1),
            "dgrad_th": lambda: gemm.conv3x3_halo(dy, wt, dgrad=True, base_mode=mode),
        }
        if gemm.conv_two_cta_ok(n, hw, hw, cout):
            cands["fprop_tc2"] = lambda: gemm.conv_fprop_nhwc(x, wt, 1, 1, stats=st, two_cta=True)
            cands["dgrad_tc2"] = lambda: gemm.conv_dgrad_nhwc(dy, wt, 1, two_cta=True)
        if gemm.halo_ok(n, hw, hw, cin, cout, 3, 3, 1, 1, pair=True):
            cands["fprop_th2"] = lambda: gemm.conv3x3_halo(x, wt, stats=st, pair=True, base_mode=mode)
            cands["dgrad_th2"] = lambda: gemm.conv3x3_halo(dy, wt, dgrad=True, pair=True, base_mode=mode)
        cands["wgrad_cudnn"] = lambda: torch.ops.aten.convolution_backward(dy, x, wt, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False])
        cands["wgrad_tc"] = lambda: gemm.conv_wgrad_nhwc(x, dy, wt.shape, 1, 1)
        if not os.environ.get("SHIPYARD_HALO_CHECK_BASIC"):      # every variant (all validated on hardware in round 2)
            cands["wgrad_th"] = lambda: gemm.conv3x3_wgrad_halo(x, dy)
            if cout == 64:
                cands["fprop_th_alt"] = lambda: gemm.conv3x3_halo(x, wt, stats=st, base_mode=mode, epi_alt=True)
                cands["dgrad_th_alt"] = lambda: gemm.conv3x3_halo(dy, wt, dgrad=True, base_mode=mode, epi_alt=True)
                cands["fprop_th2_64"] = lambda: gemm.conv3x3_halo(x, wt, stats=st, pair=True, block_n=64, base_mode=mode)
                cands["fprop_th2_64_alt"] = lambda: gemm.conv3x3_halo(x, wt, stats=st, pair=True, block_n=64, base_mode=mode, epi_alt=True)
            if cout == 128 and hw == 28:
                cands["fprop_th2_ws"] = lambda: gemm.conv3x3_halo(x, wt, stats=st, pair=True, base_mode=mode, weights_stationary=True)
                cands["dgrad_th2_ws"] = lambda: gemm.conv3x3_halo(dy, wt, dgrad=True, pair=True, base_mode=mode, weights_stationary=True)
        only = [v for v in os.environ.get("HALO_ONLY", "").split(",") if v]
        if only:                                   # one process per unverified variant (a device trap is sticky): keep the library rows + the named ones
            cands = {k: fn for k, fn in cands.items() if k.endswith("_cudnn") or any(k.endswith("_" + v) or k == v for v in only)}
        for k, fn in cands.items():
            try:
                us = t_us(fn)
                row[k] = {"us": round(us, 1), "tflops": round(flops / us / 1e6, 1)}
            except Exception as e:  # noqa: BLE001
                row[k] = {"error": str(e)[:120]}
                print(json.dumps(row), flush=True)
                return
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    what, mode = sys.argv[1], int(sys.argv[2])
    (numerics if what == "numerics" else timing)(mode)
