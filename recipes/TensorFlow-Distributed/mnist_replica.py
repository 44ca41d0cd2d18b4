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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hidden_units", type=int, default=100)
    ap.add_argument("--batch_size", type=int, default=100)
    ap.add_argument("--learning_rate", type=float, default=0.01)
    ap.add_argument("--train_steps", type=int, default=10000)
    ap.add_argument("--sync_replicas", action="store_true", default=True)
    ap.add_argument("--cuda_graph", action="store_true")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    gpu = os.environ.get("SHIPYARD_GPU", os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available() and gpu not in ("", "-1")
    dev_index = int(gpu) % max(1, torch.cuda.device_count()) if use_cuda else None
    if use_cuda:
        torch.cuda.set_device(dev_index)
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
    t0 = time.time()
    for i in range(a.train_steps):
        if graph is not None:
            graph.replay()
        else:
            loss = step()
    if graph is not None:
        loss = static_loss
    sync()
    dt = time.time() - t0
    comm.check_status()
    if rank == 0:
        print(json.dumps({"steps_per_sec": round(a.train_steps / dt, 1), "training_elapsed_s": round(dt, 3), "world": world,
                          "final_loss": round(float(loss), 4), "gradient_bytes": n * 4, "transport": comm.transport, "cuda_graph": graph is not None}), flush=True)
    comm.close()


if __name__ == "__main__":
    main()
