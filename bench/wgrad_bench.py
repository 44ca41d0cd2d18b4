#!/usr/bin/env python3
"""Split-K MN-major tcgen05 wgrad / MN-major-B dgrad kernels vs cuBLAS on the ResNet-50 1x1-conv shapes (batch 256).
CUDA-event timing, L2 flush between iterations; roofline from MEASURED_PEAKS.json (bytes: both operands once + output)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batch_shipyard_b200.ops import gemm  # noqa: E402

# (pixels, Cin, Cout)
SHAPES = [(802816, 64, 64), (802816, 64, 256), (802816, 256, 64), (200704, 256, 128), (200704, 128, 512), (200704, 512, 128),
          (50176, 512, 256), (50176, 256, 1024), (50176, 1024, 256), (12544, 1024, 512), (12544, 512, 2048), (12544, 2048, 512), (256, 2048, 1000)]


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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    os.makedirs("gpurun_out", exist_ok=True)
    out_f = open("gpurun_out/wgrad_bench.jsonl", "w")
    for (p, ci, co) in SHAPES:
        x = torch.randn(p, ci, device="cuda").to(torch.bfloat16); dy = torch.randn(p, co, device="cuda").to(torch.bfloat16)
        w = torch.randn(co, ci, device="cuda").to(torch.bfloat16)
        dw = torch.zeros(co, ci, dtype=torch.bfloat16, device="cuda"); dx = torch.empty(p, ci, dtype=torch.bfloat16, device="cuda")
        t_cub_w = timeit(lambda: torch.matmul(dy.t(), x, out=dw), 8, flush)
        t_sy_w = timeit(lambda: gemm.gemm_nt_wgrad(x, dy, out=dw, accumulate=False), 8, flush)
        t_cub_d = timeit(lambda: torch.matmul(dy, w, out=dx), 8, flush)
        t_sy_d = timeit(lambda: gemm.gemm_nn(dy, w, out=dx), 8, flush)
        fl = 2.0 * p * ci * co
        roof_w = max(fl / (peak * 1e12), 2.0 * (p * ci + p * co + ci * co) / (hbm * 1e9)) * 1e3
        row = {"pixels": p, "cin": ci, "cout": co, "wgrad_cublas_ms": round(t_cub_w, 4), "wgrad_sy_ms": round(t_sy_w, 4),
               "wgrad_speedup": round(t_cub_w / t_sy_w, 3), "wgrad_frac_of_roofline_measured": round(roof_w / t_sy_w, 3),
               "dgrad_cublas_ms": round(t_cub_d, 4), "dgrad_sy_ms": round(t_sy_d, 4), "dgrad_speedup": round(t_cub_d / t_sy_d, 3),
               "dgrad_frac_of_roofline_measured": round(roof_w / t_sy_d, 3)}
        print(json.dumps(row), flush=True); out_f.write(json.dumps(row) + "\n")


if __name__ == "__main__":
    main()
