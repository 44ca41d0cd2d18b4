"""ResNet stem (dense 4x4 convolution over the 16-channel s2d input, batch 256): tcgen05 kernels (native/gemm/stem_s2d.inc) vs cuDNN.
`--one fprop|wgrad` launches the kernel a few times for ncu."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from batch_shipyard_b200.ops import gemm


def t_us(fn, iters=20):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    n, hp = int(os.environ.get("STEM_BATCH", "256")), 115
    x = (torch.randn(n, 16, hp, hp, device="cuda") * 0.7).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 16, 4, 4, device="cuda") * 0.08).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = (torch.randn(n, 64, hp - 3, hp - 3, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    stats = torch.zeros(128, dtype=torch.float32, device="cuda")
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        for _ in range(3):
            gemm.stem_s2d_fprop(x, w, stats=stats) if sys.argv[2] == "fprop" else gemm.stem_s2d_wgrad(x, dy)
        torch.cuda.synchronize()
        return
    torch.backends.cudnn.benchmark = True
    in_b, out_b = x.numel() * 2, dy.numel() * 2
    row = {"batch": n, "roofline_us_fprop": round((in_b + out_b) / 6.58e12 * 1e6, 1), "roofline_us_wgrad": round((in_b + out_b) / 6.58e12 * 1e6, 1)}
    row["fprop_cudnn_us"] = round(t_us(lambda: F.conv2d(x, w)), 1)
    row["fprop_tc_us"] = round(t_us(lambda: gemm.stem_s2d_fprop(x, w)), 1)
    row["fprop_tc_stats_us"] = round(t_us(lambda: gemm.stem_s2d_fprop(x, w, stats=stats)), 1)
    row["wgrad_cudnn_us"] = round(t_us(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1, (False, True, False))), 1)
    row["wgrad_tc_us"] = round(t_us(lambda: gemm.stem_s2d_wgrad(x, dy)), 1)
    row["fprop_hbm_tb_s"] = round((in_b + out_b) / row["fprop_tc_us"] / 1e6, 2)
    row["wgrad_hbm_tb_s"] = round((in_b + out_b) / row["wgrad_tc_us"] / 1e6, 2)
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
