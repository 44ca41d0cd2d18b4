#!/usr/bin/env python3
"""Fused BN(+res)(+ReLU) kernels on ResNet-50 batch-256 shapes vs the measured HBM copy bandwidth.

Cold timing: a ring of buffer sets larger than 2x L2 is rotated so no call sees its operands in the 126 MB L2.
Bytes counted: fwd = x read twice (stats + apply) [+res] + out write + mask; bwd = (dout + x) read twice + dx [+dres] + mask x2.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from batch_shipyard_b200.ops import fused  # noqa: E402

SHAPES = [(112, 64), (56, 64), (56, 256), (28, 128), (28, 512), (14, 256), (14, 1024), (7, 512), (7, 2048)]


def timeit(fn, sets, iters):
    for i in range(3):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * iters)]
    for i in range(iters):
        ev[2 * i].record(); fn(sets[i % len(sets)]); ev[2 * i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(iters))
    return ts[len(ts) // 2] * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--bps3", default="", help="stats,fwd,bwd blocks per SM")
    ap.add_argument("--big-only", action="store_true", help="only the layers large enough not to be host-launch bound")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = fused.load()
    if a.blocks_per_sm:
        lib.sy_ops_set_bn_blocks_per_sm(a.blocks_per_sm)
    if a.bps3:
        lib.sy_ops_set_bn_blocks_per_sm3(*[int(v) for v in a.bps3.split(",")])
    peak = 6583.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    tot = {"fwd": 0.0, "bwd": 0.0, "fwd_ideal": 0.0, "bwd_ideal": 0.0}
    for hw, c in SHAPES:
        if a.big_only and hw * hw * c * a.batch * 2 < 100e6:
            continue
        n = a.batch
        numel = n * hw * hw * c
        nbytes = numel * 2
        nsets = max(2, int(300e6 // (nbytes * 3)) + 1)
        sets = []
        for _ in range(nsets):
            x = torch.randn(n, hw, hw, c, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2)
            res = torch.randn_like(x) if a.res else None
            dout = torch.randn_like(x)
            g = torch.nn.Parameter(torch.rand(c, device=dev, dtype=torch.bfloat16) + 0.5)
            b = torch.nn.Parameter(torch.zeros(c, device=dev, dtype=torch.bfloat16))
            sets.append((x, res, dout, g, b))
        outs = {}

        def fwd(s):
            x, res, dout, g, b = s
            xr = x.detach().requires_grad_(True)
            outs["y"] = fused.fused_bn_act(xr, g, b, res, None, None, relu=True)
            outs["xr"] = xr

        def fwdbwd(s):
            fwd(s)
            outs["y"].backward(s[2])

        t_f = timeit(fwd, sets, a.iters)
        t_fb = timeit(fwdbwd, sets, a.iters)
        t_b = t_fb - t_f
        mask_b = numel // 8
        by_f = nbytes * (3 + (1 if a.res else 0)) + mask_b
        by_b = nbytes * (5 + (1 if a.res else 0)) + 2 * mask_b
        tot["fwd"] += t_f; tot["bwd"] += t_b
        tot["fwd_ideal"] += by_f / peak / 1e3; tot["bwd_ideal"] += by_b / peak / 1e3
        print(json.dumps({"hw": hw, "c": c, "MB": round(nbytes / 1e6, 1), "fwd_us": round(t_f, 1), "fwd_gbs": round(by_f / t_f / 1e3, 1),
                          "fwd_frac_of_measured_hbm": round(by_f / t_f / 1e3 / peak, 3), "bwd_us": round(t_b, 1),
                          "bwd_gbs": round(by_b / t_b / 1e3, 1), "bwd_frac_of_measured_hbm": round(by_b / t_b / 1e3 / peak, 3)}), flush=True)
        del sets
        torch.cuda.empty_cache()
    print(json.dumps({"sum_fwd_us": round(tot["fwd"], 1), "sum_bwd_us": round(tot["bwd"], 1), "ideal_fwd_us": round(tot["fwd_ideal"], 1),
                      "ideal_bwd_us": round(tot["bwd_ideal"], 1), "hbm_gbs_measured": peak, "blocks_per_sm": a.bps3 or a.blocks_per_sm or "default"}))


if __name__ == "__main__":
    main()
