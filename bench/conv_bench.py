#!/usr/bin/env python3
"""TMA-im2col implicit-GEMM convolution (tcgen05) vs cuDNN on the ResNet-50 3x3 / strided shapes at batch 256.
CUDA events, L2 flush between iterations, median; roofline from MEASURED_PEAKS.json."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batch_shipyard_b200.ops import gemm  # noqa: E402

# (cin, hw, cout, k, stride)
SHAPES = [(64, 56, 64, 3, 1), (128, 56, 128, 3, 2), (128, 28, 128, 3, 1), (256, 28, 256, 3, 2), (256, 14, 256, 3, 1), (512, 14, 512, 3, 2),
          (512, 7, 512, 3, 1), (256, 56, 512, 1, 2), (512, 28, 1024, 1, 2), (1024, 14, 2048, 1, 2)]


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops", 1590.0)); hbm = float(peaks.get("hbm_gbs", 6650.0))
    torch.backends.cudnn.benchmark = True
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    os.makedirs("gpurun_out", exist_ok=True)
    out_f = open("gpurun_out/conv_bench.jsonl", "w")
    n = 256
    for (cin, hw, cout, k, stride) in SHAPES:
        pad = k // 2
        x = torch.randn(n, cin, hw, hw, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(cout, cin, k, k, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        p = (hw + 2 * pad - k) // stride + 1
        stats = torch.zeros(2 * cout, dtype=torch.float32, device="cuda")
        t_cudnn = timeit(lambda: F.conv2d(x, w, stride=stride, padding=pad), 8, flush)
        t_sy = timeit(lambda: gemm.conv_fprop_nhwc(x, w, stride, pad), 8, flush)
        t_sy_stats = timeit(lambda: gemm.conv_fprop_nhwc(x, w, stride, pad, stats=stats), 8, flush)
        t_2cta = None
        if cout % 128 == 0 and (n * p * p) % 256 == 0:
            t_2cta = timeit(lambda: gemm.conv_fprop_nhwc(x, w, stride, pad, stats=stats, two_cta=True), 8, flush)
        fl = 2.0 * n * p * p * cout * cin * k * k
        byts = 2.0 * (x.numel() + w.numel() + n * p * p * cout)
        roof = max(fl / (peak * 1e12), byts / (hbm * 1e9)) * 1e3
        row = {"cin": cin, "hw": hw, "cout": cout, "k": k, "stride": stride, "fprop_cudnn_ms": round(t_cudnn, 4), "fprop_sy_ms": round(t_sy, 4),
               "fprop_sy_stats_ms": round(t_sy_stats, 4), "fprop_sy_2cta_stats_ms": None if t_2cta is None else round(t_2cta, 4), "fprop_speedup": round(t_cudnn / t_sy, 3), "fprop_tflops": round(fl / t_sy / 1e9, 1),
               "fprop_frac_of_roofline_measured": round(roof / t_sy, 3)}
        dy = torch.randn(n, cout, p, p, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)

        def cudnn_bwd(dx, dw):
            return torch.ops.aten.convolution_backward(dy, x, w, None, [stride, stride], [pad, pad], [1, 1], False, [0, 0], 1, [dx, dw, False])

        if stride == 1 and k == 3:
            t_cd = timeit(lambda: cudnn_bwd(True, False), 8, flush)
            t_sd = timeit(lambda: gemm.conv_dgrad_nhwc(dy, w, pad), 8, flush)
            row.update({"dgrad_cudnn_ms": round(t_cd, 4), "dgrad_sy_ms": round(t_sd, 4), "dgrad_speedup": round(t_cd / t_sd, 3)})
        buf = torch.zeros(cout, k, k, cin, dtype=torch.bfloat16, device="cuda")
        t_cw = timeit(lambda: cudnn_bwd(False, True), 8, flush)
        t_sw = timeit(lambda: gemm.conv_wgrad_nhwc(x, dy, w.shape, stride, pad, out=buf, accumulate=True), 8, flush)
        row.update({"wgrad_cudnn_ms": round(t_cw, 4), "wgrad_sy_ms": round(t_sw, 4), "wgrad_speedup": round(t_cw / t_sw, 3)})
        print(json.dumps(row), flush=True); out_f.write(json.dumps(row) + "\n")


if __name__ == "__main__":
    main()
