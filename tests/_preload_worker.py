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
# all_to_all: PyTorch spells it ncclGroupStart; ncclSend/ncclRecv per peer; ncclGroupEnd -> ONE shipyard all-to-all kernel
for per in (256, 1 << 18):
    a_in = torch.arange(world * per, dtype=torch.float32, device=dev) + 1000.0 * rank
    a_out = torch.empty_like(a_in)
    dist.all_to_all_single(a_out, a_in)
    want = torch.cat([torch.arange(rank * per, (rank + 1) * per, dtype=torch.float32, device=dev) + 1000.0 * r for r in range(world)])
    ok &= bool((a_out == want).all())
# uneven all_to_all must be forwarded (and still be correct)
splits = [1 + (r + rank) % 3 for r in range(world)]
u_in = torch.full((sum(splits),), float(rank), device=dev)
out_splits = [1 + (rank + r) % 3 for r in range(world)]
u_out = torch.empty(sum(out_splits), device=dev)
dist.all_to_all_single(u_out, u_in, out_splits, splits)
ok &= bool((u_out == torch.cat([torch.full((out_splits[r],), float(r), device=dev) for r in range(world)])).all())
# a gradient-bucket-sized and a larger-than-staging all-reduce on plain cudaMalloc buffers (chunked through the staging halves)
for n in (25 << 18, 80 << 20):
    big = torch.full((n,), float(rank + 1), dtype=torch.bfloat16 if n > (30 << 20) else torch.float32, device=dev)
    dist.all_reduce(big)
    ok &= bool((big[:: max(1, n // 4096)].float() == world * (world + 1) / 2).all()) and float(big[-1]) == world * (world + 1) / 2
del big
# two more communicators of the SAME world size, first used in a different order on odd and even ranks: matching is by content
# (session token agreed through the real library at creation), not by call order
g1 = dist.new_group(list(range(world))); g2 = dist.new_group(list(range(world)))
dist.barrier(group=g1); dist.barrier(group=g2)        # NCCL creates a communicator at its first collective: same order everywhere
t1 = torch.full((4096,), 1.0 + rank, device=dev); t2 = torch.full((4096,), 10.0 * (1 + rank), device=dev)
w1 = dist.all_reduce(t1 if rank % 2 == 0 else t2, group=g1 if rank % 2 == 0 else g2, async_op=True)
w2 = dist.all_reduce(t2 if rank % 2 == 0 else t1, group=g2 if rank % 2 == 0 else g1, async_op=True)
w1.wait(); w2.wait()
ok &= bool((t1 == world * (world + 1) / 2).all()) and bool((t2 == 10.0 * world * (world + 1) / 2).all())
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["LD_PRELOAD"].split(":")[0])
lib.shipyard_preload_hits.restype = ctypes.c_ulonglong; lib.shipyard_preload_forwards.restype = ctypes.c_ulonglong
hits, fwd = lib.shipyard_preload_hits(), lib.shipyard_preload_forwards()
dist.barrier(); dist.destroy_process_group()
print(f"rank {rank} ok={ok} hits={hits} forwards={fwd}", flush=True)
sys.exit(0 if ok and hits >= 24 else 1)
