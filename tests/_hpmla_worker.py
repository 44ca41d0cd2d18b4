"""One rank of the SymSGD check: with the exact (full-rank) combiner, composing the ranks' local runs must reproduce SEQUENTIAL SGD over
shard 0, shard 1, ... far better than parameter averaging does, and every rank must end with the same model."""
import argparse
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from batch_shipyard_b200.ops.coll import Communicator  # noqa: E402

spec = importlib.util.spec_from_file_location("supersgd", os.path.join(ROOT, "recipes", "HPMLA-CPU-OpenMPI", "supersgd.py"))
supersgd = importlib.util.module_from_spec(spec)
spec.loader.exec_module(supersgd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, required=True)
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--session", required=True)
    ap.add_argument("--device", type=int, default=-1)
    a = ap.parse_args()
    dev_index = None if a.device < 0 else a.device
    if dev_index is not None:
        torch.cuda.set_device(dev_index)
    comm = Communicator(a.rank, a.world, a.session, dev_index, heap_bytes=96 << 20)
    dev = comm.torch_device
    dim, n, lr, batch = 24, 600, 0.05, 20
    shards = [supersgd.synthetic_shard(r, n, dim, dev) for r in range(a.world)]
    x, y = shards[a.rank]
    # reference: sequential SGD over shard 0, then shard 1, ... from w = 0 (plain torch, every rank computes it redundantly)
    seq = supersgd.SymSGD(Communicator(0, 1, a.session + f"-seq{a.rank}", None, heap_bytes=48 << 20), dim, lr, dim, 1e9, batch)
    seq.w = seq.w.to(dev); seq.A = seq.A.to(dev)
    for xs, ys in shards:
        seq.w, _ = seq.local_round(xs, ys)
    # SymSGD: one round, exact combiner (k = dim -> A = I)
    m = supersgd.SymSGD(comm, dim, lr, dim, 1e9, batch)
    w_l, n_l = m.local_round(x, y)
    locals_ = [supersgd.SymSGD(Communicator(0, 1, a.session + f"-loc{a.rank}-{r}", None, heap_bytes=48 << 20), dim, lr, dim, 1e9, batch) for r in range(a.world)]
    avg = torch.zeros(dim, device=dev)
    for r, lm in enumerate(locals_):
        lm.w = lm.w.to(dev); lm.A = lm.A.to(dev)
        avg += lm.local_round(*shards[r])[0] / a.world
    m.combine(w_l, n_l)
    err_sym = float((m.w - seq.w).norm() / seq.w.norm())
    err_avg = float((avg - seq.w).norm() / seq.w.norm())
    assert (a.world == 1 and err_sym == 0.0) or (err_sym < 0.25 * err_avg and err_sym < 0.1), (err_sym, err_avg)
    # every rank holds the same composed model
    ref = m.w.clone()
    if a.world > 1:
        comm.broadcast(ref, root=0)
        assert torch.equal(ref, m.w), "ranks disagree on the combined model"
    # low-rank projection: still a better estimate of the sequential result than averaging, and the threshold fallback averages
    m2 = supersgd.SymSGD(comm, dim, lr, 16, 1e9, batch)
    w2, n2 = m2.local_round(x, y)
    m2.combine(w2, n2)
    assert a.world == 1 or float((m2.w - seq.w).norm() / seq.w.norm()) < err_avg * 1.05
    m3 = supersgd.SymSGD(comm, dim, lr, 16, 0.0, batch)                # threshold 0: always falls back to averaging
    w3, n3 = m3.local_round(x, y)
    m3.combine(w3, n3)
    if a.world > 1:
        assert m3.fallbacks == 1 and float((m3.w - avg).norm()) < 1e-5 * float(avg.norm()) + 1e-6
    loss, acc = m.evaluate(x, y)
    assert acc > 0.8 and loss < 0.69
    print(f"rank {a.rank} symsgd: composed-vs-sequential {err_sym:.4f}, averaged-vs-sequential {err_avg:.4f}, accuracy {acc:.3f} OK")
    comm.close()


if __name__ == "__main__":
    main()
