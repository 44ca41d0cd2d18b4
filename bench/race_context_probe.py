"""Why does the dispatcher's race see the 64-column GEMMs at ~83 us when the stand-alone probe measures ~40 us?
Times gemm_nn / gemm_tn (M = 802816, N = K = 64) and cuDNN's dgrad with ops.conv._time in a growing context:
  A clean process   B after creating a Communicator (symmetric heap)   C after cuDNN benchmark-mode convolutions
  D inside ops.conv._autotune itself                                    E with 40 GB of other live allocations."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from batch_shipyard_b200.ops import conv, gemm


def measure(tag):
    n, c, hw = 256, 64, 56
    m = n * hw * hw
    x = (torch.randn(n, c, hw, hw, device="cuda") * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn_like(x)
    w4 = (torch.randn(c, c, 1, 1, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = w4.view(c, c)
    dy2 = dy.permute(0, 2, 3, 1).reshape(m, c)
    row = {"ctx": tag}
    row["nn_time"] = round(conv._time(lambda: gemm.gemm_nn(dy2, w)), 1)
    row["tn_time"] = round(conv._time(lambda: gemm.gemm_tn(dy2, w)), 1)
    row["dgrad_tc_time"] = round(conv._time(lambda: conv._dgrad_tc(dy, x, w4, 1, 0)), 1)
    row["cudnn_time"] = round(conv._time(lambda: conv._cudnn_bwd(dy, x, w4, 1, 0, True, False)), 1)
    # plain stream loop, no graph
    for name, fn in (("nn_loop", lambda: gemm.gemm_nn(dy2, w)), ("cudnn_loop", lambda: conv._cudnn_bwd(dy, x, w4, 1, 0, True, False))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); e1.synchronize()
        row[name] = round(e0.elapsed_time(e1) * 1e3 / 20, 1)
    print(json.dumps(row), flush=True)
    return x, w4


def main():
    measure("A clean")
    from batch_shipyard_b200.ops.coll import Communicator
    comm = Communicator(0, 1, device=0, heap_bytes=1 << 30)
    measure("B communicator")
    torch.backends.cudnn.benchmark = True
    xs = torch.randn(64, 64, 56, 56, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ws = torch.randn(64, 64, 3, 3, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    F.conv2d(xs, ws, padding=1)
    x, w4 = measure("C cudnn benchmark mode")
    conv.set_mode("auto")
    plan = conv._autotune(x, w4, 1)
    print(json.dumps({"ctx": "D autotune", "timings": plan.timings_us}), flush=True)
    # G: the weight operand lives in the communicator's symmetric heap (cuMemCreate memory), as the trainer's flat parameters do
    n, c, hw = 256, 64, 56
    m = n * hw * hw
    dy2 = (torch.randn(m, c, device="cuda") * 0.5).to(torch.bfloat16)
    wh = comm.alloc((c, c), torch.bfloat16); wh.copy_((torch.randn(c, c, device="cuda") * 0.05).to(torch.bfloat16))
    wt = wh.clone()
    oh = comm.alloc((m, c), torch.bfloat16)
    print(json.dumps({"ctx": "G weights in symmetric heap", "nn_heap_w": round(conv._time(lambda: gemm.gemm_nn(dy2, wh)), 1),
                      "nn_torch_w": round(conv._time(lambda: gemm.gemm_nn(dy2, wt)), 1),
                      "tn_heap_w": round(conv._time(lambda: gemm.gemm_tn(dy2, wh)), 1),
                      "tn_heap_out": round(conv._time(lambda: gemm.gemm_tn(dy2, wt, out=oh)), 1)}), flush=True)
    hog = [torch.empty(1 << 30, dtype=torch.uint8, device="cuda") for _ in range(40)]
    measure("E 40 GB live")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        measure("F on a side stream")
    del hog
    comm.close()


if __name__ == "__main__":
    main()
