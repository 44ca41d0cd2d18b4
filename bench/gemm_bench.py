#!/usr/bin/env python3
"""tcgen05 GEMM vs cuBLAS (torch.matmul) on ResNet-50 1x1-conv / FC shapes + large squares.
CUDA-event timing, L2 flush between iterations, reports TFLOP/s and fraction of the measured cuBLAS peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batch_shipyard_b200.ops import gemm  # noqa: E402

SHAPES = [(802816, 64, 64), (802816, 256, 64), (802816, 64, 256), (200704, 512, 128), (200704, 128, 512), (50176, 1024, 256),
          (50176, 256, 1024), (12544, 2048, 512), (12544, 512, 2048), (256, 1000, 2048), (8192, 8192, 8192), (4096, 4096, 4096),
          (16384, 2048, 2048)]


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
    rows = []
    for (m, n, k) in SHAPES:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16); b = torch.randn(n, k, device="cuda").to(torch.bfloat16)
        out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        stats = torch.zeros(2 * n, dtype=torch.float32, device="cuda")
        iters = 5 if m * n * k > 1e11 else 10
        t_cublas = timeit(lambda: torch.matmul(a, b.t(), out=out), iters, flush)
        best = (None, 1e9)
        for bn in (64, 128, 256):
            if bn > max(64, n):
                continue
            t = timeit(lambda: gemm.gemm_tn(a, b, out=out, block_n=bn), iters, flush)
            if t < best[1]:
                best = (bn, t)
        t_stats = timeit(lambda: gemm.gemm_tn(a, b, out=out, stats=stats, block_n=best[0]), iters, flush)
        t_2cta = t_2cta_stats = None
        if gemm.two_cta_ok(m, n):
            t_2cta = timeit(lambda: gemm.gemm_tn(a, b, out=out, two_cta=True), iters, flush)
            t_2cta_stats = timeit(lambda: gemm.gemm_tn(a, b, out=out, stats=stats, two_cta=True), iters, flush)
        fl = 2.0 * m * n * k
        byts = 2.0 * (m * k + n * k + m * n)
        roof_ms = max(fl / (peak * 1e12), byts / (hbm * 1e9)) * 1e3
        rows.append({"m": m, "n": n, "k": k, "cublas_ms": round(t_cublas, 4), "sy_ms": round(best[1], 4), "sy_block_n": best[0],
                     "sy_stats_ms": round(t_stats, 4), "sy_2cta_ms": None if t_2cta is None else round(t_2cta, 4),
                     "sy_2cta_stats_ms": None if t_2cta_stats is None else round(t_2cta_stats, 4),
                     "sy_2cta_tflops": None if t_2cta is None else round(2.0 * m * n * k / t_2cta / 1e9, 1), "sy_tflops": round(fl / best[1] / 1e9, 1), "cublas_tflops": round(fl / t_cublas / 1e9, 1),
                     "roofline_ms": round(roof_ms, 4), "sy_frac_of_roofline_measured": round(roof_ms / best[1], 3),
                     "speedup_vs_cublas": round(t_cublas / best[1], 3)})
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_bench.jsonl", "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
