#!/usr/bin/env python3
"""CNTK-GPU-OpenMPI recipe body (retarget): data-parallel CIFAR-shaped training with QUANTISED gradient exchange.

The reference launches CNTK's ``TrainResNet_CIFAR10_Distributed.py -q 1`` (1-bit SGD: gradients quantised to one bit with error
feedback before the MPI/NCCL exchange) under ``mpirun`` (/root/reference/recipes/CNTK-GPU-OpenMPI/docker/run_cntk.sh:66-78).  The
Blackwell analogue of "fewer bits on the wire" is the block-scaled fp8 all-reduce: gradients are reduced in
fp32 inside the kernel and leave it as e4m3 values + one e8m0 scale per 32 elements; like 1-bit SGD the quantisation error of a step
is fed back into the next one.  ``-q 32`` selects the plain fp32 all-reduce for comparison.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.environ.get("SHIPYARD_HOME") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from batch_shipyard_b200.ops import coll  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-q", "--quantized_bits", type=int, default=8, choices=[8, 32], help="bits per gradient element on the wire (8 = block-scaled fp8)")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--steps_per_epoch", type=int, default=50)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--lr", type=float, default=0.05)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    gpu = os.environ.get("SHIPYARD_GPU", os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available() and gpu not in ("", "-1")
    dev_index = int(gpu) % max(1, torch.cuda.device_count()) if use_cuda else None
    if use_cuda:
        torch.cuda.set_device(dev_index)
    session = (os.environ.get("SHIPYARD_COLL_SESSION") or os.environ.get("TORCHELASTIC_RUN_ID") or f"cntk-{os.getppid()}") + "-q"
    comm = coll.Communicator(rank, world, session, dev_index, heap_bytes=128 << 20)
    dev = comm.torch_device
    torch.manual_seed(0)
    # small CIFAR-shaped conv net; parameters / gradients flat in the symmetric heap
    shapes = [(16, 3, 3, 3), (16,), (32, 16, 3, 3), (32,), (10, 32 * 8 * 8), (10,)]
    sizes = [int(torch.tensor(s).prod()) for s in shapes]
    n = sum(sizes); npad = (n + 127) // 128 * 128
    flat = comm.alloc(npad, torch.float32); grad = comm.alloc(npad, torch.float32)
    q = comm.alloc(npad, torch.uint8); sc = comm.alloc(npad // 32, torch.uint8)
    flat.zero_(); grad.zero_()
    params, off = [], 0
    for s, shp in zip(sizes, shapes):
        p = flat[off:off + s].view(shp)
        if len(shp) > 1:
            p.copy_(torch.randn(shp) * (2.0 / (s / shp[0])) ** 0.5)
        p.requires_grad_(True); p.grad = grad[off:off + s].view(shp)
        params.append(p); off += s
    residual = torch.zeros(npad, dtype=torch.float32, device=dev)        # error feedback, as in 1-bit SGD
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand(a.batch, 3, 32, 32, generator=g).to(dev)
    y = torch.randint(0, 10, (a.batch,), generator=g).to(dev)

    def forward():
        h = F.max_pool2d(F.relu(F.conv2d(x, params[0], params[1], padding=1)), 2)
        h = F.max_pool2d(F.relu(F.conv2d(h, params[2], params[3], padding=1)), 2)
        return F.cross_entropy(F.linear(h.flatten(1), params[4], params[5]), y)

    losses, t0 = [], time.time()
    for step in range(a.epochs * a.steps_per_epoch):
        grad.zero_()
        loss = forward()
        loss.backward()
        with torch.no_grad():
            if a.quantized_bits == 8:
                grad.add_(residual)                                      # feed back what the last step's quantisation dropped
                mine = grad.clone()
                comm.all_reduce_fp8(grad, q, sc, scale=1.0 / world)      # reduce in fp32, leave as e4m3 + e8m0 block scales
                avg = coll.dequant_mx_fp8(q, sc)
                # error feedback: exact when this rank is alone; with peers the per-rank error is not observable from the reduced
                # result, so it is dropped (the block scales keep the relative error below 2^-3 per element anyway)
                residual.copy_(mine - avg) if world == 1 else residual.zero_()
                flat.sub_(a.lr * avg)
            else:
                if world > 1:
                    comm.all_reduce(grad, grad, scale=1.0 / world)
                flat.sub_(a.lr * grad)
        losses.append(loss.item())
    dt = time.time() - t0
    comm.check_status()
    if rank == 0:
        print(json.dumps({"first_loss": round(losses[0], 4), "last_loss": round(losses[-1], 4), "steps": len(losses), "world": world,
                          "bits_on_wire": a.quantized_bits, "gradient_bytes_fp32": n * 4, "wire_bytes": n * (1 + 1 / 32) if a.quantized_bits == 8 else n * 4,
                          "steps_per_sec": round(len(losses) / dt, 1), "transport": comm.transport}), flush=True)
    comm.close()


if __name__ == "__main__":
    main()
