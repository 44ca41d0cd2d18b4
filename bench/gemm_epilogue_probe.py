"""Probe for the short-K GEMM epilogue cost (NEXT.md "FIRST THING TO MEASURE IN ROUND 2").

Times the ResNet-50 1x1 layers as plain GEMMs (M = 256*H*W pixels, N = Cout, K = Cin) through gemm_tn (1-CTA and CTA-pair) and
reports device time, achieved HBM bandwidth and the implied cost per 128x64 epilogue chunk per warpgroup.  Run it twice:
    python bench/gemm_epilogue_probe.py                                  (TMA-store epilogue)
    SHIPYARD_GEMM_DIRECT_STORE=1 python bench/gemm_epilogue_probe.py     (st.global epilogue, kDirect)
    SHIPYARD_GEMM_DIRECT_STORE=1 SHIPYARD_GEMM_EPI_ALT=1 python bench/gemm_epilogue_probe.py   (+ alternate-tile epilogue on 64-column tiles)
The switch is read once per process by the library, hence two processes.  cuBLAS (torch.matmul) is the reference column.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from batch_shipyard_b200.ops import gemm

SHAPES = [  # pixels per image side, Cin (K), Cout (N)
    (56, 64, 64), (56, 64, 256), (56, 256, 64), (56, 256, 128), (28, 128, 512), (28, 512, 128), (28, 512, 256),
    (14, 256, 1024), (14, 1024, 256), (14, 1024, 512), (7, 512, 2048), (7, 2048, 512)]


def t_us(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def one(idx: int, two_cta: bool, stats: bool) -> None:
    """A handful of launches of one shape for `ncu` (bench/gpu_round2_ncu.sh): no timing, no warm-up loops."""
    hw, k, n = SHAPES[idx]
    m = 256 * hw * hw
    a = (torch.randn(m, k, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
    st = torch.zeros(2 * n, dtype=torch.float32, device="cuda") if stats else None
    for _ in range(3):
        gemm.gemm_tn(a, b, stats=st, two_cta=two_cta and gemm.two_cta_ok(m, n))
    torch.cuda.synchronize()
    print("probe one:", m, n, k, "two_cta" if two_cta else "1cta", "stats" if stats else "")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        return one(int(sys.argv[2]), "2cta" in sys.argv, "stats" in sys.argv)
    direct = os.environ.get("SHIPYARD_GEMM_DIRECT_STORE", "0") + ("+alt" if os.environ.get("SHIPYARD_GEMM_EPI_ALT") else "")
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for hw, k, n in SHAPES:
        m = 256 * hw * hw
        a = (torch.randn(m, k, device="cuda") * 0.5).to(torch.bfloat16)
        b = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
        st = torch.zeros(2 * n, dtype=torch.float32, device="cuda")
        row = {"direct_store": direct, "m": m, "n": n, "k": k}
        byts = 2.0 * (m * k + m * n + n * k)
        row["roofline_us"] = round(byts / 6.2e12 * 1e6, 1)
        row["cublas_us"] = round(t_us(lambda: torch.matmul(a, b.t())), 1)
        cands = {"tc": lambda: gemm.gemm_tn(a, b), "tc_stats": lambda: gemm.gemm_tn(a, b, stats=st)}
        if gemm.two_cta_ok(m, n):
            cands["tc2"] = lambda: gemm.gemm_tn(a, b, two_cta=True)
            cands["tc2_stats"] = lambda: gemm.gemm_tn(a, b, stats=st, two_cta=True)
        for name, fn in cands.items():
            us = t_us(fn)
            pair = name.startswith("tc2")
            bn = (256 if n % 256 == 0 else 128) if pair else (256 if n > 128 else (128 if n > 64 else 64))
            tiles = (m // (256 if pair else 128)) * ((n + bn - 1) // bn)
            workers = sms // 2 if pair else sms
            chunks_per_group = max(1, bn // 64 // 2)            # two epilogue warpgroups share the 64-column chunks of a tile
            per_chunk = us / (tiles / workers) / chunks_per_group
            row[name] = {"us": round(us, 1), "hbm_tb_s": round(byts / us / 1e6, 2), "us_per_chunk_per_group": round(per_chunk, 2)}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
