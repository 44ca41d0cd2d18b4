"""Stock-PyTorch arms of the headline benchmark, measured IN THE SAME PROCESS AND LEASE as the shipyard arm.

This module imports nothing from the shipyard package: torchvision's resnet50, torch.distributed (NCCL), cuDNN / cuBLAS — the stack
the reference's PyTorch-GPU recipe launches inside its container (/root/reference/recipes/PyTorch-GPU/config/jobs.yaml:1-8,
config.yaml:4-5).  Two flavours, same metric / model / batch / synthetic data shape as the shipyard arm:

  stock-eager   fp32 parameters under bf16 autocast, channels_last, DistributedDataParallel (N > 1), torch.optim.SGD, eager launch;
                the end-to-end loop copies every batch H2D synchronously and reads the loss with a host sync (what a plain training
                script does).
  stock-tuned   what a careful user gets out of stock PyTorch: bf16 parameters and activations, channels_last, one flat gradient
                buffer all-reduced with ONE NCCL call (AVG), foreach SGD, the whole step captured in a CUDA graph (NCCL capture
                included when it works, otherwise forward + backward captured and the exchange eager); the end-to-end loop
                double-buffers pinned uint8 batches on a copy stream exactly like the shipyard arm and reads the loss one step late.

Both are device-timed with CUDA events between barriers, max over ranks, with nvidia-smi clocks sampled during the timed region.
"""
from __future__ import annotations

import gc

import torch
import torch.distributed as dist
import torch.nn.functional as F

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def _barrier(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(x, world, dev):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _timed(step, steps, world, dev):
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record(); e1.synchronize()
    _barrier(world)
    return _max_over_ranks(e0.elapsed_time(e1), world, dev) / steps


def _normalise(dx_u8, out, mean, std):
    """uint8 NHWC -> normalised NCHW-logical / channels_last tensor, written into the static step input."""
    out.copy_(((dx_u8.permute(0, 3, 1, 2).to(torch.float32) / 255.0 - mean) / std))


def run_eager(batch, steps, warmup, rank, world, local, sampler_factory=None):
    import torchvision
    dev = torch.device("cuda", local)
    torch.manual_seed(1234)
    torch.backends.cudnn.benchmark = True
    model = torchvision.models.resnet50(weights=None).to(dev).to(memory_format=torch.channels_last)
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True) if world > 1 else model
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(batch, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (batch,), device=dev)
    hx = torch.randint(0, 256, (batch, 224, 224, 3), dtype=torch.uint8).pin_memory()
    hy = torch.randint(0, 1000, (batch,), dtype=torch.int64).pin_memory()
    mean = torch.tensor(MEAN, device=dev).view(1, 3, 1, 1); std = torch.tensor(STD, device=dev).view(1, 3, 1, 1)

    def step(xx, yy):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(net(xx), yy)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(3, warmup)):
        step(x, y)
    sampler = sampler_factory() if (sampler_factory and rank == 0) else None
    if sampler:
        sampler.start()
    ms = _timed(lambda i: step(x, y), steps, world, dev)
    clocks = sampler.stop() if sampler else {}

    def e2e_step(i):
        dx = hx.to(dev, non_blocking=True); dy = hy.to(dev, non_blocking=True)
        xx = ((dx.permute(0, 3, 1, 2).float() / 255.0 - mean) / std).contiguous(memory_format=torch.channels_last)
        float(step(xx, dy))

    for i in range(3):
        e2e_step(i)
    ms_e2e = _timed(e2e_step, steps, world, dev)
    out = {"flavour": "stock-eager", "what": "torchvision resnet50, fp32 params + bf16 autocast, channels_last, DDP/NCCL, SGD, eager",
           "ms_per_step": round(ms, 3), "value": round(batch * world / (ms / 1e3), 2), "unit": "images/sec",
           "e2e": {"value": round(batch * world / (ms_e2e / 1e3), 2), "ms_per_step": round(ms_e2e, 3),
                   "h2d_bytes_per_step": hx.numel() + hy.numel() * 8, "d2h_bytes_per_step": 4, "input_path": "synchronous"},
           "clocks": clocks}
    del net, model, opt, x, y
    gc.collect(); torch.cuda.empty_cache()
    return out


def run_tuned(batch, steps, warmup, rank, world, local, sampler_factory=None):
    import torchvision
    dev = torch.device("cuda", local)
    torch.manual_seed(1234)
    torch.backends.cudnn.benchmark = True
    model = torchvision.models.resnet50(weights=None).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    params = [p for p in model.parameters() if p.requires_grad]
    flat_g = torch.zeros(sum(p.numel() for p in params), dtype=torch.bfloat16, device=dev)
    off = 0
    for p in params:                                  # static gradients: views of ONE flat buffer -> one NCCL call per step
        n = p.numel()
        g = flat_g[off:off + n].view(p.shape)
        if p.dim() == 4:
            g = flat_g[off:off + n].view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)   # channels_last like the weight
        p.grad = g
        off += n
    opt = torch.optim.SGD(params, lr=0.1, momentum=0.9, weight_decay=1e-4, foreach=True)
    x = torch.zeros(batch, 3, 224, 224, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.zeros(batch, dtype=torch.int64, device=dev)
    x.copy_(torch.randn(batch, 3, 224, 224, device=dev)); y.copy_(torch.randint(0, 1000, (batch,), device=dev))
    loss_buf = torch.zeros((), dtype=torch.float32, device=dev)

    def fwd_bwd():
        loss = F.cross_entropy(model(x).float(), y)
        loss.backward()                               # accumulates into the static flat views (zeroed at the end of the step)
        loss_buf.copy_(loss.detach())

    def exchange_update():
        if world > 1:
            dist.all_reduce(flat_g, op=dist.ReduceOp.AVG)
        opt.step()
        flat_g.zero_()

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(max(3, warmup)):
            fwd_bwd(); exchange_update()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    import os
    mode = "whole step in one CUDA graph (NCCL all-reduce captured)" if world > 1 else "whole step in one CUDA graph"
    g_full = g_fb = None
    try:
        if world > 1 and not os.environ.get("SHIPYARD_BASELINE_CAPTURE_NCCL"):
            # capturing the NCCL call is opt-in: a failed capture poisons the stream for the rest of the process, and this arm
            # shares the process with the headline measurement
            raise RuntimeError("NCCL capture not requested")
        g_full = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_full):
            fwd_bwd(); exchange_update()
        torch.cuda.synchronize()
        g_full.replay(); torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001 - NCCL capture unsupported / not requested: capture forward + backward only
        g_full = None
        torch.cuda.synchronize()
        mode = f"forward+backward in a CUDA graph, all-reduce + SGD eager ({type(e).__name__})"
        try:
            flat_g.zero_()
            g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fb):
                fwd_bwd()
        except Exception as e2:  # noqa: BLE001
            g_fb = None
            mode = f"eager ({type(e2).__name__})"

    def step(_i=0):
        if g_full is not None:
            g_full.replay()
        elif g_fb is not None:
            g_fb.replay(); exchange_update()
        else:
            fwd_bwd(); exchange_update()

    for _ in range(max(3, warmup)):
        step()
    sampler = sampler_factory() if (sampler_factory and rank == 0) else None
    if sampler:
        sampler.start()
    ms = _timed(step, steps, world, dev)
    clocks = sampler.stop() if sampler else {}

    # end to end: pinned uint8 batches, double-buffered H2D on a copy stream, convert on the GPU, loss read one step late
    depth = 2
    hx = [torch.randint(0, 256, (batch, 224, 224, 3), dtype=torch.uint8).pin_memory() for _ in range(depth)]
    hy = [torch.randint(0, 1000, (batch,), dtype=torch.int64).pin_memory() for _ in range(depth)]
    dx = [torch.empty((batch, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(depth)]
    dy = [torch.empty((batch,), dtype=torch.int64, device=dev) for _ in range(depth)]
    hloss = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(depth)]
    copied = [torch.cuda.Event() for _ in range(depth)]; consumed = [torch.cuda.Event() for _ in range(depth)]
    ready = [torch.cuda.Event() for _ in range(depth)]
    copy_stream = torch.cuda.Stream(dev)
    issued = [False] * depth
    mean = torch.tensor(MEAN, device=dev).view(1, 3, 1, 1); std = torch.tensor(STD, device=dev).view(1, 3, 1, 1)

    def prefetch(s):
        with torch.cuda.stream(copy_stream):
            if issued[s]:
                copy_stream.wait_event(consumed[s])
            dx[s].copy_(hx[s], non_blocking=True); dy[s].copy_(hy[s], non_blocking=True)
            copied[s].record(copy_stream)
        issued[s] = True

    def run(s):
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(copied[s])
        _normalise(dx[s], x, mean, std)
        y.copy_(dy[s], non_blocking=True)
        consumed[s].record(cur)
        step()
        hloss[s].copy_(loss_buf, non_blocking=True)
        ready[s].record(cur)

    def read(s):
        ready[s].synchronize()
        return float(hloss[s])

    for i in range(3):
        prefetch(i % 2); run(i % 2); read(i % 2)

    def e2e_step(i):
        s = i % 2
        if i == 0:
            prefetch(s)
        run(s)
        if i + 1 < steps:
            prefetch((i + 1) % 2)
        if i > 0:
            read((i - 1) % 2)

    ms_e2e = _timed(e2e_step, steps, world, dev)
    read((steps - 1) % 2)
    out = {"flavour": "stock-tuned",
           "what": "torchvision resnet50, bf16 params, channels_last, flat gradient + ONE NCCL all-reduce(AVG), foreach SGD; " + mode,
           "ms_per_step": round(ms, 3), "value": round(batch * world / (ms / 1e3), 2), "unit": "images/sec",
           "e2e": {"value": round(batch * world / (ms_e2e / 1e3), 2), "ms_per_step": round(ms_e2e, 3),
                   "h2d_bytes_per_step": hx[0].numel() + hy[0].numel() * 8, "d2h_bytes_per_step": 4,
                   "input_path": "double-buffered copy stream"},
           "clocks": clocks}
    del g_full, g_fb, model, opt, params, flat_g, x, y
    gc.collect(); torch.cuda.empty_cache()
    return out


def run_compiled(batch, steps, warmup, rank, world, local, sampler_factory=None):
    """stock-compiled (on by default since the end of round 2, SHIPYARD_BASELINE_COMPILE=0 skips it; inductor needs ~1 minute per
    process): the stock recipe with torch.compile on the model, so the BatchNorm / ReLU / residual elementwise chains are fused by the
    tracing compiler.  The strongest stock arm measured on this box (9 331 img/s on one B200 vs 4 983 for stock-tuned)."""
    import time
    import torchvision
    dev = torch.device("cuda", local)
    torch.manual_seed(1234)
    torch.backends.cudnn.benchmark = True
    model = torchvision.models.resnet50(weights=None).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True) if world > 1 else model
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4, foreach=True)
    cnet = torch.compile(net)
    x = torch.randn(batch, 3, 224, 224, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (batch,), device=dev)

    def step(_i=0):
        opt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(cnet(x).float(), y)
        loss.backward()
        opt.step()
        return loss

    t0 = time.time()
    for _ in range(max(3, warmup)):
        step()
    torch.cuda.synchronize()
    compile_s = time.time() - t0
    sampler = sampler_factory() if (sampler_factory and rank == 0) else None
    if sampler:
        sampler.start()
    ms = _timed(step, steps, world, dev)
    clocks = sampler.stop() if sampler else {}
    out = {"flavour": "stock-compiled", "what": "stock-tuned recipe + torch.compile(model) (inductor), DDP/NCCL, foreach SGD",
           "ms_per_step": round(ms, 3), "value": round(batch * world / (ms / 1e3), 2), "unit": "images/sec",
           "compile_and_warmup_s": round(compile_s, 1), "clocks": clocks}
    del cnet, net, model, opt
    gc.collect(); torch.cuda.empty_cache()
    return out


def run_both(batch, steps, warmup, rank, world, local, sampler_factory=None) -> dict:
    import os
    out = {}
    arms = [("stock_eager", run_eager), ("stock_tuned", run_tuned)]
    # torch.compile arm: on by default on one GPU; at N > 1 every rank would run inductor at the same time (minutes of host time on a
    # shared box), so there it is opt-in (SHIPYARD_BASELINE_COMPILE=1) and bench/baseline_measured.json carries the scaled 1-GPU number
    if os.environ.get("SHIPYARD_BASELINE_COMPILE", "1" if world == 1 else "0") not in ("0", "", "off", "false"):
        arms.append(("stock_compiled", run_compiled))
    for name, fn in arms:
        try:
            out[name] = fn(batch, steps, warmup, rank, world, local, sampler_factory)
        except Exception as e:  # noqa: BLE001 - a failing baseline arm must not take the headline measurement down
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            gc.collect(); torch.cuda.empty_cache()
    return out
