"""torch.distributed (NCCL backend) all_reduce / all_gather / reduce_scatter / broadcast under LD_PRELOAD of the shim."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
dev = torch.device("cuda", rank)
ok = True
for n in (5, 4096, 1 << 20, (3 << 20) + 7):
    for dt in (torch.float32, torch.bfloat16):
        x = torch.full((n,), float(rank + 1), dtype=dt, device=dev)
        dist.all_reduce(x)
        ok &= bool((x.float() == world * (world + 1) / 2).all())
        y = torch.full((n,), float(rank + 1), dtype=dt, device=dev)
        dist.all_reduce(y, op=dist.ReduceOp.AVG)
        ok &= bool(((y.float() - (world + 1) / 2).abs() < 1e-2).all())
g_in = torch.full((1000,), float(rank), device=dev); g_out = torch.empty(1000 * world, device=dev)
dist.all_gather_into_tensor(g_out, g_in)
ok &= all(bool((g_out[r * 1000:(r + 1) * 1000] == r).all()) for r in range(world))
rs_in = torch.arange(1024 * world, dtype=torch.float32, device=dev); rs_out = torch.empty(1024, device=dev)
dist.reduce_scatter_tensor(rs_out, rs_in)
ok &= bool((rs_out == world * torch.arange(rank * 1024, (rank + 1) * 1024, device=dev)).all())
b = torch.full((777,), float(rank), device=dev); dist.broadcast(b, src=world - 1)
ok &= bool((b == world - 1).all())
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["LD_PRELOAD"].split(":")[0])
lib.shipyard_preload_hits.restype = ctypes.c_ulonglong; lib.shipyard_preload_forwards.restype = ctypes.c_ulonglong
hits, fwd = lib.shipyard_preload_hits(), lib.shipyard_preload_forwards()
dist.barrier(); dist.destroy_process_group()
print(f"rank {rank} ok={ok} hits={hits} forwards={fwd}", flush=True)
sys.exit(0 if ok and hits >= 18 else 1)
