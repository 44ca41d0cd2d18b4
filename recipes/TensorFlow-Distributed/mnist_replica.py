#!/usr/bin/env python3
"""TensorFlow-Distributed recipe body, retargeted from a gRPC parameter server to NVSwitch all-reduce.

The reference runs TF-1.2 ``mnist_replica.py`` with one parameter server and N workers
(/root/reference/recipes/TensorFlow-Distributed/docker/gpu/launcher.sh:21-77,
mnist_replica.py:59-80,111-181): 784 -> 100 (ReLU) -> 10 softmax, batch 100, Adam(0.01),
``train_steps``, optional ``sync_replicas``.  Same model and flags here; the parameter server is replaced
by a synchronous all-reduce of the 79,510-parameter gradient (318 KB fp32: the latency-bound regime,
served by the one-shot push kernel) fused with the 1/N averaging AND the Adam update of every replica (one launch per step).
TensorFlow is not installed in this image, so the model is PyTorch; data is synthetic MNIST-shaped.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.environ.get("SHIPYARD_HOME") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from batch_shipyard_b200.ops.coll import Communicator  # noqa: E402


def run_nccl_arm(a, rank, world, dev_index, use_cuda) -> dict:
    """Stock PyTorch arm (no shipyard code on the path): the same 784 -> h -> 10 MLP, batch and Adam(lr) with the gradient averaged by
    torch.distributed's NCCL all-reduce — what a TensorFlow/PyTorch user gets when the parameter server is swapped for all-reduce
    with the library stack.  Two flavours: eager (one flat all-reduce + torch.optim.Adam per step) and the whole step in a CUDA graph
    (capturable Adam, NCCL captured) when the build supports it."""
    import torch.distributed as dist
    dev = torch.device("cuda", dev_index) if use_cuda else torch.device("cpu")
    created = False
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl" if use_cuda else "gloo", rank=rank, world_size=world, **({"device_id": dev} if use_cuda else {}))
        created = True
    torch.manual_seed(0)
    h = a.hidden_units
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand(a.batch_size, 784, generator=g).to(dev)
    y = torch.randint(0, 10, (a.batch_size,), generator=g).to(dev)
    out = {"impl": "nccl", "world": world, "flavours": {}}

    def build(capturable):
        torch.manual_seed(0)
        n = 784 * h + h + h * 10 + 10
        flat = torch.zeros(n, device=dev); gflat = torch.zeros(n, device=dev)
        flat.copy_(torch.cat([torch.randn(784 * h) * (1.0 / 28), torch.zeros(h), torch.randn(h * 10) * (h ** -0.5), torch.zeros(10)]))
        ps, off = [], 0
        for cnt, shape in zip([784 * h, h, h * 10, 10], [(h, 784), (h,), (10, h), (10,)]):
            p = flat[off:off + cnt].view(shape).requires_grad_(True)
            p.grad = gflat[off:off + cnt].view(shape)
            ps.append(p); off += cnt
        opt = torch.optim.Adam(ps, lr=a.learning_rate, capturable=capturable and use_cuda, foreach=True)

        def step():
            loss = F.cross_entropy(F.linear(torch.relu(F.linear(x, ps[0], ps[1])), ps[2], ps[3]), y)
            loss.backward()
            if world > 1:
                dist.all_reduce(gflat, op=dist.ReduceOp.AVG if use_cuda else dist.ReduceOp.SUM)
                if not use_cuda:
                    gflat.div_(world)
            opt.step()
            gflat.zero_()
            return loss
        return step

    def timed(run, steps):
        if use_cuda:
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if use_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.time()
        for _ in range(steps):
            run()
        if use_cuda:
            e1.record(); e1.synchronize()
            dt = e0.elapsed_time(e1) / 1e3
        else:
            dt = time.time() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t)
        return dt

    steps = a.train_steps
    step = build(False)
    for _ in range(20):
        step()
    dt = timed(step, steps)
    out["flavours"]["eager"] = {"steps_per_sec": round(steps / dt, 1), "training_elapsed_s": round(dt, 3)}
    if use_cuda:
        try:
            step = build(True)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(20):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                step()
            torch.cuda.synchronize()
            for _ in range(20):
                gr.replay()
            dt = timed(gr.replay, steps)
            out["flavours"]["cuda_graph"] = {"steps_per_sec": round(steps / dt, 1), "training_elapsed_s": round(dt, 3)}
        except Exception as e:  # noqa: BLE001 - capture of NCCL / capturable Adam unsupported: the eager flavour stands
            out["flavours"]["cuda_graph"] = {"error": f"{type(e).__name__}: {e}"[:200]}
            torch.cuda.synchronize()
    if created and a.impl == "nccl":
        dist.barrier(); dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hidden_units", type=int, default=100)
    ap.add_argument("--batch_size", type=int, default=100)
    ap.add_argument("--learning_rate", type=float, default=0.01)
    ap.add_argument("--train_steps", type=int, default=10000)
    ap.add_argument("--sync_replicas", action="store_true", default=True)
    ap.add_argument("--cuda_graph", action="store_true")
    ap.add_argument("--impl", default="shipyard", choices=["shipyard", "nccl", "both"],
                    help="nccl: the same model on stock PyTorch (torch.distributed NCCL all-reduce + torch.optim.Adam); both: NCCL arm, then "
                         "the shipyard arm, in one process, with the ratio")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    gpu = os.environ.get("SHIPYARD_GPU", os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available() and gpu not in ("", "-1")
    dev_index = int(gpu) % max(1, torch.cuda.device_count()) if use_cuda else None
    if use_cuda:
        torch.cuda.set_device(dev_index)
    nccl_res = None
    if a.impl in ("nccl", "both"):
        nccl_res = run_nccl_arm(a, rank, world, dev_index, use_cuda)
        if a.impl == "nccl":
            if rank == 0:
                print(json.dumps(nccl_res), flush=True)
            return
    comm = Communicator(rank, world, os.environ.get("SHIPYARD_COLL_SESSION", f"tfdist-{os.getppid()}") + "-mlp", dev_index, heap_bytes=128 << 20)
    dev = comm.torch_device
    torch.manual_seed(0)
    h = a.hidden_units
    sizes = [784 * h, h, h * 10, 10]
    n = sum(sizes)
    npad = (n + 7) // 8 * 8
    flat = comm.alloc(npad, torch.float32)          # parameters and gradients live in the symmetric heap: zero-copy all-reduce
    grad = comm.alloc(npad, torch.float32)
    flat.zero_(); grad.zero_()
    init = torch.cat([torch.randn(784 * h) * (1.0 / 28), torch.zeros(h), torch.randn(h * 10) * (h ** -0.5), torch.zeros(10)])
    flat[:n].copy_(init)
    views, gviews, off = [], [], 0
    for s, shape in zip(sizes, [(h, 784), (h,), (10, h), (10,)]):
        p = flat[off:off + s].view(shape).requires_grad_(True)
        p.grad = grad[off:off + s].view(shape)
        views.append(p); off += s
    # Adam state for the fused kernel: one launch = gradient all-reduce (1/N fused) + Adam on this replica + gradient zeroing
    adam_m = torch.zeros(npad, dtype=torch.float32, device=dev)
    adam_v = torch.zeros(npad, dtype=torch.float32, device=dev)
    hyper = torch.tensor([a.learning_rate, 0.9, 0.999, 1e-8, 0.0], dtype=torch.float32, device=dev)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand(a.batch_size, 784, generator=g).to(dev)
    y = torch.randint(0, 10, (a.batch_size,), generator=g).to(dev)

    def step():
        logits = F.linear(torch.relu(F.linear(x, views[0], views[1])), views[2], views[3])
        loss = F.cross_entropy(logits, y)
        loss.backward()                                  # accumulates into `grad` (zeroed by the fused kernel)
        comm.fused_allreduce_adam(grad, flat, adam_m, adam_v, hyper, scale=(1.0 / world) if a.sync_replicas else 1.0)
        return loss

    for _ in range(20):
        step()
    sync = (lambda: torch.cuda.synchronize()) if use_cuda else (lambda: None)
    graph = None
    static_loss = torch.zeros((), device=dev)
    if use_cuda and (a.cuda_graph or os.environ.get("SHIPYARD_TF_GRAPH", "1") != "0"):
        # forward + backward + ONE fused all-reduce/Adam kernel captured as a CUDA graph: a training step is one launch
        sync()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss.copy_(step().detach())
    sync(); comm.barrier(); sync()
    if use_cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.time()
    for i in range(a.train_steps):
        if graph is not None:
            graph.replay()
        else:
            loss = step()
    if graph is not None:
        loss = static_loss
    if use_cuda:
        e1.record(); e1.synchronize()
    sync()
    dt = time.time() - t0
    if use_cuda:                                          # device-timed, max over ranks
        t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float32, device=dev)
        ts = comm.alloc(8, torch.float32); ts.zero_(); ts[:1].copy_(t); sync(); comm.barrier(); sync()
        if world > 1:
            comm.all_reduce(ts, ts, op="max")
        sync()
        dt = float(ts[0])
    comm.check_status()
    if rank == 0:
        out = {"steps_per_sec": round(a.train_steps / dt, 1), "training_elapsed_s": round(dt, 3), "world": world, "impl": "shipyard",
               "final_loss": round(float(loss), 4), "gradient_bytes": n * 4, "transport": comm.transport, "cuda_graph": graph is not None,
               "timing": "cuda events, max over ranks" if use_cuda else "wall clock"}
        if nccl_res is not None:
            out["nccl_same_run"] = nccl_res
            best = max(v["steps_per_sec"] for v in nccl_res["flavours"].values() if "steps_per_sec" in v)
            out["vs_best_nccl_flavour"] = round(out["steps_per_sec"] / best, 3)
        print(json.dumps(out), flush=True)
    comm.close()
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass


if __name__ == "__main__":
    main()
