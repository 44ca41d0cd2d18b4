#!/usr/bin/env python3
"""HPCG recipe body (retarget): one rank per GPU, z-slab decomposition, NVLink halo exchange.

The reference recipe runs Intel's prebuilt ``xhpcg_skx --n=256 --t=120`` through ``mpirun -hosts ... -perhost 1``
over InfiniBand (/root/reference/recipes/HPCG-Infiniband-IntelMPI/config/docker/jobs.yaml:5-20).  Here the
task runner starts the ranks; ``--n`` is the local cube edge and ``--t`` the timed seconds, as in xhpcg.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.environ.get("SHIPYARD_HOME") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch  # noqa: E402

from batch_shipyard_b200.models.hpcg import HPCG, HPCGValidityError  # noqa: E402
from batch_shipyard_b200.ops.coll import Communicator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--size", dest="n", type=int, default=256, help="local grid edge (nx=ny=nz); --size for launchers that abbreviate-match --n")
    ap.add_argument("--t", "--seconds", dest="t", type=float, default=30.0, help="timed seconds")
    ap.add_argument("--levels", type=int, default=4)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    gpu = os.environ.get("SHIPYARD_GPU", os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available() and gpu not in ("", "-1")
    dev_index = int(gpu) % max(1, torch.cuda.device_count()) if use_cuda else None
    if use_cuda:
        torch.cuda.set_device(dev_index)
    downscaled = None
    if not use_cuda and not os.environ.get("SHIPYARD_CPU_FULL_SIZE") and (a.n > 32 or a.t > 5):
        # virtual (CPU) slots are a functional mode: the GPU-sized problem of the recipe would run for hours on CPU threads
        downscaled = {"requested_n": a.n, "requested_t": a.t}
        a.n, a.t, a.levels = min(a.n, 32), min(a.t, 5.0), min(a.levels, 3)
    session = (os.environ.get("SHIPYARD_COLL_SESSION") or os.environ.get("TORCHELASTIC_RUN_ID") or f"hpcg-{os.getppid()}") + "-hpcg"
    comm = Communicator(rank, world, session, dev_index, heap_bytes=256 << 20)
    try:
        res = HPCG(comm, a.n, a.n, a.n, levels=a.levels).benchmark(seconds=a.t)
    except HPCGValidityError as e:
        # validity gate (symmetry of A and of the preconditioner, optimised path == eager fp64 path): no GFLOP/s for an invalid run
        if rank == 0:
            print(json.dumps({"valid": False, "error": str(e), "world": world, "local_grid": [a.n, a.n, a.n]}), flush=True)
        comm.close()
        sys.exit(1)
    if rank == 0:
        if downscaled:
            res["downscaled_for_cpu"] = downscaled
        print(json.dumps({k: (round(v, 6) if isinstance(v, float) and k != "residual_reduction" else v) for k, v in res.items()}), flush=True)
    comm.close()


if __name__ == "__main__":
    main()
