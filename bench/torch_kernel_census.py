"""Which kernels of a training step are NOT ours, and which aten op (with which shapes / strides) launched them?

`profiles/step_breakdown.md` shows ~3 ms per step in `at::elementwise_kernel` / `at::vectorized_elementwise_kernel` (15 % of the
step): residual-gradient adds, but also 21 launches per step of the NON-vectorised copy kernel at ~88 us each, i.e. a layout
conversion of a large activation once per block.  This tool runs a few EAGER steps (no CUDA graph, so every launch is attributed)
of the bench trainer under torch.profiler and prints, per aten op + input shapes + strides, the CUDA time per step, sorted.

    python bench/torch_kernel_census.py [--steps 2] [--batch 256] [--top 40]         (1 GPU; profiler numbers are for attribution only)
"""
import argparse
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    from torch.profiler import ProfilerActivity, profile
    from batch_shipyard_b200.models.resnet import resnet50
    from batch_shipyard_b200.ops.coll import Communicator
    from batch_shipyard_b200.parallel.ddp import FusedDataParallelTrainer
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    comm = Communicator(0, 1, session=f"census-{os.getpid()}", device=0, heap_bytes=1 << 30)
    tr = FusedDataParallelTrainer(resnet50(), comm, (a.batch, 3, 224, 224), 1000, lr=0.1, momentum=0.9, weight_decay=1e-4, use_graph=False)
    g = torch.Generator(device="cpu").manual_seed(7)
    tr.load_images_u8(torch.randint(0, 256, (a.batch, 224, 224, 3), dtype=torch.uint8, generator=g).to(dev),
                      torch.randint(0, 1000, (a.batch,), generator=g).to(dev))
    tr.prepare(warmup=3)                       # autotunes the conv dispatcher; eager steps from here on
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
        for _ in range(a.steps):
            tr.step()
        torch.cuda.synchronize()
    # 1. kernels by name
    kern = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = kern[e.name[:110]]; k[0] += 1; k[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    total = sum(v[1] for v in kern.values())
    print(f"# kernels ({a.steps} steps, total {total / a.steps / 1e3:.2f} ms/step under the profiler)")
    for name, (n, us) in sorted(kern.items(), key=lambda kv: -kv[1][1])[: a.top]:
        ours = name.startswith(("k_", "gemm_bf16", "conv3x3_halo", "wgrad3x3"))
        print(f"{us / a.steps:10.1f} us/step  {n / a.steps:6.1f} launches/step  {'own ' if ours else 'LIB '} {name}")
    # 2. aten ops that launched library elementwise / copy kernels, grouped by input shapes
    print("\n# aten ops by (name, input shapes): CUDA time per step")
    rows = []
    for e in prof.key_averages(group_by_input_shape=True):
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = e.cuda_time_total
        if t <= 0 or not e.key.startswith("aten::"):
            continue
        rows.append((t / a.steps, e.count / a.steps, e.key, str(e.input_shapes)[:160]))
    for us, cnt, key, shapes in sorted(rows, reverse=True)[: a.top]:
        print(f"{us:10.1f} us/step  {cnt:6.1f} calls/step  {key:38s} {shapes}")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kernel_census.json", "w") as f:
        json.dump({"kernels": {k: v for k, v in kern.items()}, "aten": rows}, f)
    comm.close()


if __name__ == "__main__":
    main()
