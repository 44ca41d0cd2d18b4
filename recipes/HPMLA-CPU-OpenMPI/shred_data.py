#!/usr/bin/env python3
"""Shred a training file into one shard per rank for supersgd.py: ``<out_prefix>.<rank>`` (round-robin by line, so every shard sees
the same class mix).  Counterpart of the reference's Data-Shredding step, which splits the data set by node and thread count and uploads
the pieces to a blob container mounted on every node (/root/reference/recipes/HPMLA-CPU-OpenMPI/Data-Shredding/README.md:8-22); here the
pieces go to a directory every rank can read (the pool's shared directory, ``$AZ_BATCH_NODE_SHARED_DIR``), e.g. from a job-preparation
task or `shipyard data ingress`.

    shred_data.py --input train.libsvm --out-prefix $AZ_BATCH_NODE_SHARED_DIR/hpmla/train --node-count 8 [--thread-count 1]
    shred_data.py --synthetic 200000 --dim 256 --out-prefix ... --node-count 8        # writes a separable synthetic set instead
"""
import argparse
import os
import random


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", default="")
    ap.add_argument("--out-prefix", required=True)
    ap.add_argument("--node-count", type=int, required=True)
    ap.add_argument("--thread-count", type=int, default=1, help="shards per node (the reference shreds by node x thread); ranks read <prefix>.<rank>")
    ap.add_argument("--synthetic", type=int, default=0, help="generate this many examples instead of reading --input")
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    shards = a.node_count * max(1, a.thread_count)
    os.makedirs(os.path.dirname(os.path.abspath(a.out_prefix)) or ".", exist_ok=True)
    outs = [open(f"{a.out_prefix}.{i}", "w") for i in range(shards)]
    n = 0
    try:
        if a.synthetic:
            rng = random.Random(a.seed)
            w = [rng.gauss(0, 1) for _ in range(a.dim)]
            for n in range(a.synthetic):
                x = [rng.gauss(0, 1) for _ in range(a.dim)]
                label = 1 if sum(wi * xi for wi, xi in zip(w, x)) + 0.3 * rng.gauss(0, 1) > 0 else -1
                outs[n % shards].write(str(label) + " " + " ".join(f"{i}:{v:.5f}" for i, v in enumerate(x)) + "\n")
            n = a.synthetic
        else:
            with open(a.input) as f:
                for n, line in enumerate(f):
                    if line.strip():
                        outs[n % shards].write(line if line.endswith("\n") else line + "\n")
                n += 1
    finally:
        for o in outs:
            o.close()
    print(f"shredded {n} examples into {shards} shards: {a.out_prefix}.0 .. {a.out_prefix}.{shards - 1}")


if __name__ == "__main__":
    main()
