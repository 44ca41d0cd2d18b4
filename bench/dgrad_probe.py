"""1x1 data-gradient probe: dX[M, Cin] = dY[M, Cout] @ W[Cout, Cin] for every ResNet-50 1x1 layer, three ways:

    cudnn    aten.convolution_backward (input gradient only) on channels_last tensors
    nn       gemm_nn: W read in place as an MN-major tcgen05 B operand (what the dispatcher calls dgrad_tc / dgrad_tc2)
    tn_wt    gemm_tn on a pre-transposed weight W^T[Cin, Cout] (K-major B, the fprop kernel)

Inputs rotate through enough copies to exceed the 126 MB L2 (the dispatcher's race sees cold inputs too), so the numbers are HBM
numbers.  `--one <idx> <nn|tn_wt>` launches a few kernels of one shape for ncu.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from batch_shipyard_b200.ops import gemm

SHAPES = [  # H = W, Cin, Cout of the forward 1x1 convolution
    (56, 64, 64), (56, 64, 256), (56, 256, 64), (56, 256, 128), (28, 128, 512), (28, 512, 128), (28, 512, 256),
    (14, 256, 1024), (14, 1024, 256), (14, 1024, 512), (7, 512, 2048), (7, 2048, 512)]
L2_BYTES = 160 << 20


def rotating(make, nbytes):
    return [make() for _ in range(max(2, int(L2_BYTES // max(nbytes, 1)) + 1))]


def t_us(fn, n_rot, iters=24):
    for i in range(4):
        fn(i % n_rot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_rot)
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        hw, cin, cout = SHAPES[int(sys.argv[2])]
        m = 256 * hw * hw
        dy = (torch.randn(m, cout, device="cuda") * 0.5).to(torch.bfloat16)
        w = (torch.randn(cout, cin, device="cuda") * 0.05).to(torch.bfloat16)
        wt = w.t().contiguous()
        for _ in range(3):
            gemm.gemm_nn(dy, w) if sys.argv[3] == "nn" else gemm.gemm_tn(dy, wt)
        torch.cuda.synchronize()
        return
    for hw, cin, cout in SHAPES:
        n = 256
        m = n * hw * hw
        byts = 2.0 * (m * cout + m * cin + cin * cout)
        dys = rotating(lambda: (torch.randn(n, cout, hw, hw, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last),
                       2 * m * (cin + cout))
        nr = len(dys)
        x = torch.empty(n, cin, hw, hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w4 = (torch.randn(cout, cin, 1, 1, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = w4.view(cout, cin)
        wt = w.t().contiguous()
        dy2 = [d.permute(0, 2, 3, 1).reshape(m, cout) for d in dys]
        row = {"hw": hw, "cin": cin, "cout": cout, "m": m, "roofline_us": round(byts / 6.58e12 * 1e6, 1)}
        ref = torch.ops.aten.convolution_backward(dys[0], x, w4, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1, (True, False, False))[0]
        ref2 = ref.permute(0, 2, 3, 1).reshape(m, cin).float()
        cands = {
            "cudnn": lambda i: torch.ops.aten.convolution_backward(dys[i], x, w4, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1, (True, False, False)),
            "nn": lambda i: gemm.gemm_nn(dy2[i], w),
            "tn_wt": lambda i: gemm.gemm_tn(dy2[i], wt),
        }
        if gemm.two_cta_ok(m, cin):
            cands["nn2"] = lambda i: gemm.gemm_nn(dy2[i], w, two_cta=True)
            cands["tn_wt2"] = lambda i: gemm.gemm_tn(dy2[i], wt, two_cta=True)
        for name, fn in cands.items():
            if name != "cudnn":
                got = fn(0).float()
                err = float((got - ref2).norm() / ref2.norm().clamp_min(1e-6))
                if err > 2e-2:
                    row[name] = {"error": round(err, 4)}
                    continue
            us = t_us(fn, nr)
            row[name] = {"us": round(us, 1), "hbm_tb_s": round(byts / us / 1e6, 2)}
        print(json.dumps(row), flush=True)
        del dys, dy2


if __name__ == "__main__":
    main()
