#!/usr/bin/env python3
"""mpiBench-style collective sweep: our kernels vs NCCL on the same GPUs.

torchrun --nproc-per-node N bench/coll_sweep.py [--max-bytes 1G] [--ops allreduce,...]
Device-timed (CUDA events), max over ranks.  Sizes >= 1 MiB: one event pair per
iteration with an L2 flush (256 MiB write) in between; smaller sizes: batches of
back-to-back calls (latency regime; peers' data arrives over NVLink, not from L2).
Bus bandwidth uses the nccl-tests conventions (allreduce 2(N-1)/N, allgather /
reduce_scatter / alltoall (N-1)/N, broadcast 1).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batch_shipyard_b200.ops.coll import Communicator  # noqa: E402


def parse_size(s):
    s = s.strip().upper()
    m = {"K": 1 << 10, "M": 1 << 20, "G": 1 << 30}
    return int(float(s[:-1]) * m[s[-1]]) if s[-1] in m else int(s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-bytes", default="1K")
    ap.add_argument("--max-bytes", default="1G")
    ap.add_argument("--step", type=int, default=4, help="size multiplier between points")
    ap.add_argument("--ops", default="allreduce,allgather,reduce_scatter,alltoall,broadcast")
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--out", default="gpurun_out/coll_sweep.jsonl")
    ap.add_argument("--algos", default="auto,ll,oneshot,twoshot_p2p,twoshot_nvls")
    ap.add_argument("--max-blocks", default="")
    ap.add_argument("--nvls-copy", default="")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dtype = getattr(torch, a.dtype)
    esz = torch.empty((), dtype=dtype).element_size()
    lo, hi = parse_size(a.min_bytes), parse_size(a.max_bytes)
    heap = max(1 << 30, 3 * hi + (256 << 20))
    comm = Communicator(rank, world, session=f"sweep-{os.environ.get('MASTER_PORT')}-{os.getppid()}", device=local, heap_bytes=heap)
    if a.max_blocks:
        comm.set_tuning(max_blocks=int(a.max_blocks))
    if a.nvls_copy != "":
        comm.set_tuning(nvls_copy=int(a.nvls_copy), nvls_min_world=2 if int(a.nvls_copy) else 99)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sizes = []
    s = lo
    while s <= hi:
        sizes.append(s); s *= a.step
    rows = []

    def measure(fn, nbytes):
        small = nbytes < (1 << 20)
        iters = 50 if nbytes <= (64 << 10) else (20 if small else (10 if nbytes <= (64 << 20) else 5))
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        if small:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record(); e1.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
        else:
            ts = []
            for _ in range(iters):
                flush.fill_(1)
                dist.barrier()                 # ranks enter together: a barrier-free collective would otherwise be charged the host skew
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            us = ts[len(ts) // 2]              # median: one host hiccup in ten iterations must not decide the row
        t = torch.tensor([us], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    busf = {"allreduce": 2 * (world - 1) / world, "allgather": (world - 1) / world, "reduce_scatter": (world - 1) / world,
            "alltoall": (world - 1) / world, "broadcast": 1.0}
    algos = a.algos.split(",")
    for op in a.ops.split(","):
        for nbytes in sizes:
            n = max(world, nbytes // esz // world * world)     # total elements (per-rank buffer)
            nbytes = n * esz
            comm.reset_heap()
            x = comm.alloc(n, dtype); y = comm.alloc(n, dtype)
            x.normal_(); torch.cuda.synchronize(); comm.barrier(); torch.cuda.synchronize()
            xn = torch.randn(n, dtype=torch.float32, device=dev).to(dtype); yn = torch.empty_like(xn)
            res = {"op": op, "bytes": nbytes, "world": world, "dtype": a.dtype}
            if op == "allreduce":
                res["nccl_us"] = measure(lambda: dist.all_reduce(xn), nbytes)
                for al in algos:
                    if al == "ll" and nbytes > 16 << 10: continue
                    if al == "oneshot" and nbytes > 8 << 20: continue
                    if al == "twoshot_nvls" and not comm.has_multicast: continue
                    if al in ("twoshot_p2p", "twoshot_nvls") and nbytes < 4096: continue
                    res[f"sy_{al}_us"] = measure(lambda: comm.all_reduce(x, x, algo=al), nbytes)
                # fused 1/N scale + bf16->fp32 cast variant vs NCCL + separate scale/cast kernels
            elif op == "allgather":
                per = n // world
                res["nccl_us"] = measure(lambda: dist.all_gather_into_tensor(yn, xn[:per]), nbytes)
                res["sy_auto_us"] = measure(lambda: comm.all_gather(x[:per], y), nbytes)
            elif op == "reduce_scatter":
                per = n // world
                res["nccl_us"] = measure(lambda: dist.reduce_scatter_tensor(yn[:per], xn), nbytes)
                res["sy_auto_us"] = measure(lambda: comm.reduce_scatter(x, yn[:per]), nbytes)
            elif op == "alltoall":
                res["nccl_us"] = measure(lambda: dist.all_to_all_single(yn, xn), nbytes)
                res["sy_auto_us"] = measure(lambda: comm.all_to_all(x, y), nbytes)
            elif op == "broadcast":
                res["nccl_us"] = measure(lambda: dist.broadcast(xn, 0), nbytes)
                res["sy_auto_us"] = measure(lambda: comm.broadcast(x, 0), nbytes)
            best = min(v for k, v in res.items() if k.startswith("sy_"))
            res["sy_best_us"] = best
            res["speedup_vs_nccl"] = round(res["nccl_us"] / best, 3)
            res["sy_busbw_GBs"] = round(nbytes * busf[op] / best / 1e3, 2)
            res["nccl_busbw_GBs"] = round(nbytes * busf[op] / res["nccl_us"] / 1e3, 2)
            comm.check_status()
            rows.append(res)
            if rank == 0:
                print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
    if rank == 0:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    comm.close()
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
