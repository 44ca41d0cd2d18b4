#!/usr/bin/env python3
"""HPMLA recipe body (retarget): SymSGD-style distributed linear learner, one rank per GPU (or per CPU slot).

The reference recipe runs the closed `supersgd` binary of Microsoft's High Performance ML Algorithms through
``mpirun -np $nodes supersgd -l <lr> -k <rank> -mc <threshold> -e <epochs> -r <rounds> -f <prefix> -t <threads> -gl <n> -glDir <dir>``
on training data that was shredded into one file per node and thread beforehand
(/root/reference/recipes/HPMLA-CPU-OpenMPI/docker/run_parasail.sh:86, README.md:30-52, Data-Shredding/README.md).  The algorithm
family is published (SymSGD: every worker runs *sequential* SGD on its shard and also tracks a low-rank "model combiner" — how its
result would change had it started from a different model — so the workers' results can be composed as if the shards had been
processed one after the other instead of being averaged).  This body implements that scheme for logistic regression at mini-batch
granularity with the same knobs:

  -l learning rate   -k rank of the combiner projection   -m combiner threshold (bound on |M_i - I| sqrt(d / k), the estimated error of
  the projected combiner, beyond which a round falls back to plain averaging)   -e epochs   -r rounds per epoch   -f shard prefix (``<prefix>.<rank>`` written by
  shred_data.py)   -t threads (CPU torch threads)   -g log the global model every g epochs   -d directory for those logs

One round on rank i, starting from the global model w_g:  w_i <- w_g; N_i <- A (A: d x k projection shared by all ranks,
E[A A^T] = I).  For every mini-batch (X, y):  p = sigmoid(X w_i);  w_i -= l X^T (p - y) / B;  N_i -= l X^T (s * (X N_i)) / B with
s = p (1 - p)  (the Jacobian I - l X^T S X / B applied to the projected combiner).  After the round every rank all-gathers
[w_i | N_i] (one collective) and composes in rank order:  w <- w_0;  d = w - w_g;  w <- w_i + d + (N_i - A) A^T d  for i = 1..W-1
(M_i = I + (M_i - I): only the small second term is seen through the projection).
Loss / accuracy are all-reduced.  Collectives run on the shipyard kernels (NVLink) or the stub transport on CPU pools.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.environ.get("SHIPYARD_HOME") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch  # noqa: E402

from batch_shipyard_b200.ops.coll import Communicator  # noqa: E402


def load_shard(path: str, dim: int, device):
    """A shard is text, one example per line: ``label idx:val idx:val ...`` (libsvm, 0- or 1-based indices < dim) or ``label,v0,v1,...``."""
    xs, ys = [], []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if ":" in line:
                parts = line.split()
                row = torch.zeros(dim)
                for tok in parts[1:]:
                    i, v = tok.split(":")
                    row[int(i) % dim] = float(v)
                ys.append(float(parts[0]))
            else:
                vals = [float(v) for v in line.split(",")]
                ys.append(vals[0])
                row = torch.zeros(dim)
                row[:len(vals) - 1] = torch.tensor(vals[1:dim + 1])
            xs.append(row)
    x = torch.stack(xs).to(device) if xs else torch.zeros((0, dim), device=device)
    y = torch.tensor(ys, device=device)
    return x, (y > 0).float()


def synthetic_shard(rank: int, n: int, dim: int, device, seed: int = 11):
    """Linearly separable-ish data from one hidden model shared by all ranks (used when no -f prefix is given)."""
    g = torch.Generator().manual_seed(seed)
    w_true = torch.randn(dim, generator=g)
    g2 = torch.Generator().manual_seed(seed * 7919 + rank + 1)
    x = torch.randn(n, dim, generator=g2)
    y = (x @ w_true + 0.3 * torch.randn(n, generator=g2) > 0).float()
    return x.to(device), y.to(device)


class SymSGD:
    def __init__(self, comm, dim: int, lr: float, k: int, threshold: float, batch: int, seed: int = 3):
        self.comm, self.W, self.R = comm, comm.world, comm.rank
        self.dev = comm.torch_device
        self.dim, self.lr, self.k, self.thr, self.batch = dim, lr, k, threshold, batch
        if k >= dim:                                                    # full rank: the exact combiner (A = I), used by the tests
            self.k = dim
            self.A = torch.eye(dim, device=self.dev)
        else:
            g = torch.Generator().manual_seed(seed)                     # the same projection on every rank
            self.A = (torch.randn(dim, k, generator=g) / (k ** 0.5)).to(self.dev)
        self.w = torch.zeros(dim, device=self.dev)
        self.fallbacks = 0

    def local_round(self, x, y):
        w, n = self.w.clone(), self.A.clone()
        for b0 in range(0, x.shape[0], self.batch):
            xb, yb = x[b0:b0 + self.batch], y[b0:b0 + self.batch]
            p = torch.sigmoid(xb @ w)
            s = p * (1 - p)
            scale = self.lr / xb.shape[0]
            n -= scale * (xb.t() @ (s.unsqueeze(1) * (xb @ n)))          # combiner first: it uses the pre-update probabilities
            w -= scale * (xb.t() @ (p - yb))
        return w, n

    def combine(self, w_local, n_local):
        """All-gather [w_i | N_i] and compose in rank order (identical arithmetic on every rank)."""
        if self.W == 1:
            self.w = w_local
            return
        mine = torch.cat([w_local.unsqueeze(1), n_local], dim=1).contiguous()          # d x (k + 1)
        allm = torch.empty((self.W,) + tuple(mine.shape), device=self.dev)
        self.comm.all_gather(mine.view(-1), allm.view(-1))
        if self.dev.type == "cuda":
            torch.cuda.synchronize(self.dev)
        # the projection sees only (M_i - I); its error grows with |M_i - I| * sqrt(d / k).  Beyond the threshold (-m) the composed
        # model would be noisier than a plain average, so that round averages instead (exact combiner, k = d: never)
        if self.k < self.dim:
            dev = max(float((allm[i, :, 1:] - self.A).norm() / self.A.norm()) for i in range(self.W)) * (self.dim / self.k) ** 0.5
            if dev > self.thr:
                self.fallbacks += 1
                self.w = allm[:, :, 0].mean(dim=0)
                return
        w_g = self.w
        w = allm[0, :, 0].clone()
        for i in range(1, self.W):
            # M_i = I + (M_i - I) and only the small second term goes through the projection: M_i d ~= d + (N_i - A) A^T d
            d = w - w_g
            w = allm[i, :, 0] + d + (allm[i, :, 1:] - self.A) @ (self.A.t() @ d)
        self.w = w

    def evaluate(self, x, y):
        z = x @ self.w
        loss = torch.nn.functional.binary_cross_entropy_with_logits(z, y, reduction="sum")
        stat = torch.stack([loss, ((z > 0).float() == y).float().sum(), torch.tensor(float(x.shape[0]), device=self.dev)]).float()
        if self.W > 1:
            self.comm.all_reduce(stat, stat)
            if self.dev.type == "cuda":
                torch.cuda.synchronize(self.dev)
        return float(stat[0] / stat[2]), float(stat[1] / stat[2])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-l", dest="lr", type=float, default=1e-1)
    ap.add_argument("-k", dest="k", type=int, default=32)
    ap.add_argument("-m", "-mc", dest="thr", type=float, default=1e-2)
    ap.add_argument("-e", dest="epochs", type=int, default=10)
    ap.add_argument("-r", dest="rounds", type=int, default=10)
    ap.add_argument("-f", dest="prefix", default="")
    ap.add_argument("-t", dest="threads", type=int, default=1)
    ap.add_argument("-g", "-gl", dest="log_every", type=int, default=0)
    ap.add_argument("-d", "-glDir", dest="log_dir", default="")
    ap.add_argument("-b", "-bd", dest="bindir", default="", help="accepted for launch-line compatibility (the reference's binary directory)")
    ap.add_argument("-w", dest="workdir", default="", help="accepted for launch-line compatibility")
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--examples", type=int, default=20000, help="synthetic examples per rank when no -f prefix is given")
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    gpu = os.environ.get("SHIPYARD_GPU", os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available() and gpu not in ("", "-1")
    dev_index = int(gpu) % max(1, torch.cuda.device_count()) if use_cuda else None
    if use_cuda:
        torch.cuda.set_device(dev_index)
    else:
        torch.set_num_threads(max(1, a.threads))
    session = (os.environ.get("SHIPYARD_COLL_SESSION") or os.environ.get("TORCHELASTIC_RUN_ID") or f"hpmla-{os.getppid()}") + "-hpmla"
    comm = Communicator(rank, world, session, dev_index, heap_bytes=192 << 20)
    dev = comm.torch_device
    if a.prefix:
        x, y = load_shard(f"{a.prefix}.{rank}", a.dim, dev)
    else:
        x, y = synthetic_shard(rank, a.examples, a.dim, dev)
    model = SymSGD(comm, a.dim, a.lr, a.k, a.thr, a.batch)
    per_round = (x.shape[0] + a.rounds - 1) // max(1, a.rounds)
    t0 = time.time()
    history = []
    for ep in range(a.epochs):
        for r in range(a.rounds):
            xs, ys = x[r * per_round:(r + 1) * per_round], y[r * per_round:(r + 1) * per_round]
            if xs.shape[0] == 0:
                xs, ys = x[:0], y[:0]
            w_l, n_l = model.local_round(xs, ys)
            model.combine(w_l, n_l)
        loss, acc = model.evaluate(x, y)
        history.append((loss, acc))
        if rank == 0 and a.log_every and (ep + 1) % a.log_every == 0 and a.log_dir:
            os.makedirs(a.log_dir, exist_ok=True)
            with open(os.path.join(a.log_dir, f"global_model_epoch_{ep + 1}.txt"), "w") as f:
                f.write("\n".join(f"{v:.8g}" for v in model.w.tolist()) + "\n")
    if rank == 0:
        total = float(x.shape[0]) * world * a.epochs
        print(json.dumps({"algorithm": "symsgd-logistic", "world": world, "dim": a.dim, "rank_k": a.k, "epochs": a.epochs, "rounds_per_epoch": a.rounds,
                          "examples_per_rank": int(x.shape[0]), "first_epoch_loss": round(history[0][0], 5), "final_loss": round(history[-1][0], 5),
                          "final_accuracy": round(history[-1][1], 5), "averaging_fallback_rounds": model.fallbacks,
                          "examples_per_sec": round(total / max(time.time() - t0, 1e-9), 1), "transport": comm.transport}), flush=True)
    comm.close()


if __name__ == "__main__":
    main()
