"""K10: tcgen05 GEMM + all-reduce in one kernel vs fp32 reference (and timing vs cuBLAS + NCCL when asked)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batch_shipyard_b200.ops import coll, gemm  # noqa: E402


def gen(rank, m, k, salt):
    g = torch.Generator().manual_seed(100 * salt + rank)
    return (torch.randn(m, k, generator=g) * 0.25).to(torch.bfloat16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, required=True); ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--session", required=True); ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--transport", default="auto"); ap.add_argument("--bench", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(a.device)
    comm = coll.Communicator(a.rank, a.world, a.session, a.device, heap_bytes=1 << 30, transport=a.transport)
    dev = comm.torch_device
    for (m, n, kshard) in [(256, 128, 64), (1000, 520, 192), (4096, 1024, 512)]:
        A = gen(a.rank, m, kshard, 1).to(dev); B = gen(a.rank, n, kshard, 2).to(dev)
        out = comm.alloc((m, n), torch.float32)
        out.zero_(); torch.cuda.synchronize(); comm.barrier(); torch.cuda.synchronize()
        gemm.gemm_tn_allreduce(comm, A, B, out)
        torch.cuda.synchronize()
        ref = torch.zeros(m, n, dtype=torch.float64)
        for r in range(a.world):
            ref += gen(r, m, kshard, 1).double() @ gen(r, n, kshard, 2).double().t()
        err = (out.cpu().double() - ref).abs().max().item()
        assert err < 2e-3 * (kshard * a.world) ** 0.5 + 1e-3, f"K10 {m}x{n}x{kshard}: max err {err}"
        # v2: reduce-scatter / all-gather schedule, bf16 wire format, fp32 cross-rank accumulation
        out2 = comm.alloc((m, n), torch.bfloat16)
        out2.fill_(7.0); torch.cuda.synchronize(); comm.barrier(); torch.cuda.synchronize()
        for _ in range(2):                      # twice: inbox reuse / epochs
            gemm.gemm_tn_allreduce_bf16(comm, A, B, out2)
        torch.cuda.synchronize()
        ref2 = torch.zeros(m, n, dtype=torch.float64)
        for r in range(a.world):
            ref2 += (gen(r, m, kshard, 1).double() @ gen(r, n, kshard, 2).double().t()).to(torch.bfloat16).double()   # partials travel as bf16
        err2 = (out2.cpu().double() - ref2).abs().max().item()
        assert err2 < 0.02 * ref2.abs().max().item() + 0.02, f"K10v2 {m}x{n}x{kshard}: max err {err2}"
        comm.barrier(); torch.cuda.synchronize(); comm.reset_heap(); comm.__dict__.pop("_k10_inbox", None)
    line = f"rank {a.rank} transport={comm.transport} multicast={comm.has_multicast} K10 numerics OK"
    if a.bench:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29588")
        dist.init_process_group("nccl", rank=a.rank, world_size=a.world, device_id=dev)
        m, n, k = 8192, 8192, 8192 // a.world
        A = gen(a.rank, m, k, 3).to(dev); B = gen(a.rank, n, k, 4).to(dev)
        out = comm.alloc((m, n), torch.float32)
        tmp = torch.empty(m, n, dtype=torch.bfloat16, device=dev)

        def fused():
            out.zero_()
            gemm.gemm_tn_allreduce(comm, A, B, out)

        out_bf = comm.alloc((m, n), torch.bfloat16)

        def fused_v2():
            gemm.gemm_tn_allreduce_bf16(comm, A, B, out_bf)

        def baseline():
            torch.matmul(A, B.t(), out=tmp)
            dist.all_reduce(tmp)

        res = {}
        for name, fn in (("fused", fused), ("fused_v2", fused_v2), ("cublas_nccl", baseline)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record(); e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[name] = round(float(t), 4)
        line += f" | bench m=n=8192 k/rank={k}: fused multimem.red(zero+gemm+allreduce fp32 out) {res['fused']} ms, fused RS/AG bf16 {res['fused_v2']} ms, cuBLAS+NCCL(bf16) {res['cublas_nccl']} ms"
        dist.destroy_process_group()
    comm.check_status()
    comm.close()
    print(line + " OK", flush=True)


if __name__ == "__main__":
    main()
