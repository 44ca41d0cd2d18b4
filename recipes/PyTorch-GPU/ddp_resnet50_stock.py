#!/usr/bin/env python3
"""Stock PyTorch data-parallel ResNet-50: torchvision model + torch.nn.parallel.DistributedDataParallel + the NCCL backend.

This file deliberately imports NOTHING from the shipyard package: it is the "user container" of the reference's PyTorch-GPU recipe
(/root/reference/recipes/PyTorch-GPU/config/jobs.yaml:1-8 launches a stock PyTorch image) retargeted to ResNet-50.  Submitted with
`shipyard jobs add`, the task runner starts one rank per GPU and LD_PRELOADs libshipyard_preload.so, so the ncclAllReduce /
ncclBroadcast / ncclAllGather calls PyTorch issues resolve to the shipyard NVLS / P2P kernels; with SHIPYARD_COLL_DISABLE=1 the very
same job runs on stock NCCL.  Prints one JSON line (images/sec, device-timed, max over ranks) on rank 0.
"""
import argparse
import ctypes
import json
import os

import torch
import torch.distributed as dist
import torch.nn.functional as F
import torchvision


def shim_counters():
    """(collectives on shipyard kernels, forwarded to NCCL) if the preload shim is in this process, else None."""
    for path in (os.environ.get("LD_PRELOAD") or "").split(":"):
        if "shipyard_preload" in path:
            try:
                lib = ctypes.CDLL(path)
                lib.shipyard_preload_hits.restype = ctypes.c_ulonglong
                lib.shipyard_preload_forwards.restype = ctypes.c_ulonglong
                return int(lib.shipyard_preload_hits()), int(lib.shipyard_preload_forwards())
            except OSError:
                return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--bucket-mb", type=int, default=25)
    ap.add_argument("--eager-fp32", action="store_true", help="fp32 parameters under bf16 autocast instead of bf16 parameters")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("SHIPYARD_GPU", os.environ.get("LOCAL_RANK", str(rank)))) % max(1, torch.cuda.device_count())
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234)
    model = getattr(torchvision.models, a.model)(weights=None).to(dev).to(memory_format=torch.channels_last)
    if not a.eager_fp32:
        model = model.to(torch.bfloat16)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], bucket_cap_mb=a.bucket_mb, gradient_as_bucket_view=True)
    opt = torch.optim.SGD(ddp.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(a.batch, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
    x = x if a.eager_fp32 else x.to(torch.bfloat16)
    y = torch.randint(0, 1000, (a.batch,), device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.eager_fp32):
            loss = F.cross_entropy(ddp(x).float(), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record(); e1.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    counters = shim_counters()
    tot = torch.tensor(list(counters) if counters else [0, 0], device=dev, dtype=torch.int64)
    dist.all_reduce(tot)
    if rank == 0:
        print(json.dumps({"recipe": "PyTorch-GPU stock DDP", "model": a.model, "n_gpus": world, "per_gpu_batch": a.batch,
                          "ms_per_step": round(float(ms), 3), "images_per_sec": round(world * a.batch / float(ms) * 1e3, 1),
                          "loss": round(float(loss), 4), "params_dtype": "fp32+autocast" if a.eager_fp32 else "bf16",
                          "shim_loaded": counters is not None, "collectives_on_shipyard_kernels": int(tot[0]),
                          "collectives_forwarded_to_nccl": int(tot[1]),
                          "coll_disabled": bool(os.environ.get("SHIPYARD_COLL_DISABLE"))}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
