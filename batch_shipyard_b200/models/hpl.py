"""HPL-MxP (mixed-precision LINPACK) for the HPLinpack recipe retarget.

The reference recipe launches Intel's prebuilt MKL ``mp_linpack`` (fp64 LU with partial pivoting on CPUs) over a P x Q MPI grid
(/root/reference/recipes/HPLinpack-Infiniband-IntelMPI/config/docker/jobs.yaml:5-28, ``runme_intel64_prv -p $P -q $Q -b $B $PSIZE``
with the problem size picked by ``setup_hplinpack.sh`` / ``findpq.py``).  Nothing of it can run here (CPU binary in a third-party
image) and an fp64 LU is the wrong benchmark for Blackwell, whose fp64 units are a small fraction of its tensor throughput.  The
B200-native body solves the same dense system ``A x = b`` to fp64 accuracy the HPL-MxP way:

* ``A`` is the HPL-AI matrix family: uniform(-0.5, 0.5) entries with a dominant diagonal, so LU needs no pivoting; it is regenerated
  block by block from a seed whenever the fp64 residual is needed, only its fp32 working copy is stored.
* right-looking block LU, 1-D block-cyclic over the ranks by column blocks of width ``nb`` (P x Q = 1 x world): the owner factors the
  diagonal block, inverts its triangles, forms ``L21 = A21 * inv(U11)`` on the tensor cores and broadcasts ``[inv(L11); L21]`` in bf16
  over NVLink (symmetric-heap broadcast); every rank then computes ``U12 = inv(L11) * A12`` and the trailing update
  ``A22 -= L21 * U12`` with the tcgen05 bf16 GEMM (fp32 accumulation in TMEM) of ``ops.gemm``.
* the low-precision factors precondition iterative refinement in fp64: ``r = b - A x`` (fp64, all-reduce of the per-rank partial
  products), ``x += (LU)^-1 r`` with block forward / backward substitution (the owner of a column block updates the whole right-hand
  side and broadcasts it), until HPL's scaled residual ``|r|_inf / (eps * (|A|_inf |x|_inf + |b|_inf) * n)`` is below 16.

GFLOP/s uses HPL's operation count ``2/3 n^3 + 3/2 n^2`` over the whole solve (factorisation + refinement), device-timed, max over ranks.
On CPU tensors (tests: stub communicator, world 1-2) the GEMMs are emulated with the same rounding points (bf16 operands, fp32
accumulation, bf16 product).
"""
from __future__ import annotations

import time
from typing import Optional

import torch

_EPS64 = 2.0 ** -53


class HPLError(RuntimeError):
    pass


def _lu_nopivot_(d: torch.Tensor, base: int = 32) -> None:
    """In-place LU without pivoting of a square block (unit lower / upper in one array); recursive so that almost all work is GEMM."""
    n = d.shape[0]
    if n <= base:
        for i in range(n - 1):
            d[i + 1:, i] /= d[i, i]
            d[i + 1:, i + 1:] -= torch.outer(d[i + 1:, i], d[i, i + 1:])
        return
    h = n // 2
    _lu_nopivot_(d[:h, :h], base)
    # U12 = L11^-1 A12 ; L21 = A21 U11^-1
    d[:h, h:] = torch.linalg.solve_triangular(d[:h, :h], d[:h, h:], upper=False, unitriangular=True)
    d[h:, :h] = torch.linalg.solve_triangular(d[:h, :h], d[h:, :h], upper=True, left=False)
    d[h:, h:] -= d[h:, :h] @ d[:h, h:]
    _lu_nopivot_(d[h:, h:], base)


class HPLMxP:
    def __init__(self, comm, n: int, nb: int = 2048, seed: int = 42, update_chunk: int = 8192):
        if n % nb:
            raise HPLError(f"n ({n}) must be a multiple of the block size ({nb})")
        self.comm, self.n, self.nb, self.seed = comm, n, nb, seed
        self.W, self.R = comm.world, comm.rank
        self.dev = comm.torch_device
        self.cuda = self.dev.type == "cuda"
        self.nblk = n // nb
        self.mine = [j for j in range(self.nblk) if j % self.W == self.R]          # global column-block indices, ascending
        self.update_chunk = max(nb, update_chunk // nb * nb)
        self.launches0 = self._launches()
        # symmetric buffers: the bf16 panel [inv(L11); L21] and the fp32 right-hand side of the substitutions
        self.panel = comm.alloc(n * nb, torch.bfloat16).view(n, nb)
        self.vec = comm.alloc(n, torch.float32)
        self.aloc = torch.empty((n, len(self.mine) * nb), dtype=torch.float32, device=self.dev)
        self.b = self._rand(self.nblk, n, 1).view(n)                                   # fp64
        rowsum = torch.zeros(n, dtype=torch.float64, device=self.dev)
        for li, j in enumerate(self.mine):
            blk = self._a_block(j)
            rowsum += blk.abs().sum(dim=1)
            self.aloc[:, li * nb:(li + 1) * nb] = blk.float()
        self._allreduce64(rowsum)
        self.norm_a = float(rowsum.max())
        self.norm_b = float(self.b.abs().max())

    # ---- problem generation (fp64, regenerated on demand) -----------------------------------------------------------
    def _rand(self, block: int, rows: int, cols: int) -> torch.Tensor:
        g = torch.Generator(device=self.dev)
        g.manual_seed(self.seed * 1000003 + block)
        return torch.rand((rows, cols), generator=g, dtype=torch.float64, device=self.dev) - 0.5

    def _a_block(self, j: int) -> torch.Tensor:
        """Column block j of A in fp64: uniform(-0.5, 0.5) with n / 2 added on the diagonal (strictly diagonally dominant)."""
        blk = self._rand(j, self.n, self.nb)
        r0 = j * self.nb
        blk[r0:r0 + self.nb].diagonal().add_(0.5 * self.n)
        return blk

    # ---- helpers -------------------------------------------------------------------------------------------------------
    def _launches(self) -> int:
        if not self.cuda:
            return 0
        from ..ops import gemm
        return gemm.launch_count()

    def _allreduce64(self, t: torch.Tensor) -> None:
        if self.W > 1:
            self.comm.all_reduce(t, t)

    def _mm(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """bf16 [M,K] @ bf16 [K,N] -> bf16 [M,N], fp32 accumulation: the tcgen05 GEMM on the GPU, the same rounding points on the CPU."""
        if self.cuda:
            from ..ops import gemm
            m, nn = a.shape[0], b.shape[1]
            return gemm.gemm_nn(a, b, two_cta=gemm.two_cta_ok(m, nn) and m >= 256)
        return (a.float() @ b.float()).to(torch.bfloat16)

    def _first_local_after(self, k: int) -> int:
        """Number of my column blocks with global index <= k (= local index of the first trailing block)."""
        return sum(1 for j in self.mine if j <= k)

    # ---- factorisation ---------------------------------------------------------------------------------------------------
    def factor(self) -> None:
        n, nb = self.n, self.nb
        eye = torch.eye(nb, dtype=torch.float32, device=self.dev)
        for k in range(self.nblk):
            owner, r0 = k % self.W, k * nb
            m = n - r0
            if self.R == owner:
                lk = self.mine.index(k)
                pan = self.aloc[r0:, lk * nb:(lk + 1) * nb]
                d = pan[:nb]
                if self.cuda:
                    lu, _ = torch.linalg.lu_factor(d, pivot=False)
                    d.copy_(lu)
                else:
                    _lu_nopivot_(d)
                inv_l = torch.linalg.solve_triangular(d, eye, upper=False, unitriangular=True)
                inv_u = torch.linalg.solve_triangular(d, eye, upper=True)
                self.panel[r0:r0 + nb] = inv_l.to(torch.bfloat16)
                if m > nb:
                    l21 = self._mm(pan[nb:].to(torch.bfloat16), inv_u.to(torch.bfloat16).contiguous())
                    self.panel[r0 + nb:] = l21
                    pan[nb:] = l21.float()                       # the factor that is stored is the one that was used
            if self.W > 1:
                self.comm.broadcast(self.panel[r0:], root=owner)
            first = self._first_local_after(k)
            c0 = first * nb
            nc = self.aloc.shape[1] - c0
            if nc == 0:
                continue
            inv_l16, l21 = self.panel[r0:r0 + nb], self.panel[r0 + nb:]
            for cc in range(c0, c0 + nc, self.update_chunk):
                ce = min(cc + self.update_chunk, c0 + nc)
                a12 = self.aloc[r0:r0 + nb, cc:ce]
                u12 = self._mm(inv_l16, a12.to(torch.bfloat16).contiguous())
                a12.copy_(u12)
                if m > nb:
                    self.aloc[r0 + nb:, cc:ce].sub_(self._mm(l21, u12))

    # ---- (LU)^-1 r with the stored factors: fp32, block substitutions, right-hand side replicated ------------------------
    def solve_lu(self, r: torch.Tensor) -> torch.Tensor:
        n, nb = self.n, self.nb
        v = self.vec
        v.copy_(r.float())
        for k in range(self.nblk):                                # forward: L y = r (unit lower)
            owner, r0 = k % self.W, k * nb
            if self.R == owner:
                lk = self.mine.index(k)
                col = self.aloc[:, lk * nb:(lk + 1) * nb]
                yk = torch.linalg.solve_triangular(col[r0:r0 + nb], v[r0:r0 + nb].unsqueeze(1), upper=False, unitriangular=True)
                v[r0:r0 + nb] = yk.squeeze(1)
                if r0 + nb < n:
                    v[r0 + nb:] -= (col[r0 + nb:] @ yk).squeeze(1)
            if self.W > 1:
                self.comm.broadcast(v[r0:], root=owner)
        for k in range(self.nblk - 1, -1, -1):                    # backward: U x = y
            owner, r0 = k % self.W, k * nb
            if self.R == owner:
                lk = self.mine.index(k)
                col = self.aloc[:, lk * nb:(lk + 1) * nb]
                xk = torch.linalg.solve_triangular(col[r0:r0 + nb], v[r0:r0 + nb].unsqueeze(1), upper=True)
                v[r0:r0 + nb] = xk.squeeze(1)
                if r0 > 0:
                    v[:r0] -= (col[:r0] @ xk).squeeze(1)
            if self.W > 1:
                self.comm.broadcast(v[:r0 + nb], root=owner)
        return v.double()

    # ---- fp64 residual with the regenerated matrix -------------------------------------------------------------------------
    def residual(self, x: torch.Tensor) -> torch.Tensor:
        nb = self.nb
        ax = torch.zeros(self.n, dtype=torch.float64, device=self.dev)
        for j in self.mine:
            ax += self._a_block(j) @ x[j * nb:(j + 1) * nb]
        self._allreduce64(ax)
        return self.b - ax

    def scaled_residual(self, r: torch.Tensor, x: torch.Tensor) -> float:
        return float(r.abs().max()) / (_EPS64 * (self.norm_a * float(x.abs().max()) + self.norm_b) * self.n)

    # ---- the benchmark ----------------------------------------------------------------------------------------------------
    def solve(self, max_refine: int = 50, tol: float = 16.0) -> dict:
        sync = (lambda: torch.cuda.synchronize(self.dev)) if self.cuda else (lambda: None)
        self.comm.barrier(); sync()
        if self.cuda:
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
        t0 = time.perf_counter()
        self.factor()
        if self.cuda:
            e1.record()
        t1 = time.perf_counter()
        x = self.solve_lu(self.b)
        history = []
        for it in range(max_refine + 1):
            r = self.residual(x)
            history.append(self.scaled_residual(r, x))
            if history[-1] < tol or it == max_refine:
                break
            x = x + self.solve_lu(r)
        if self.cuda:
            e2.record(); sync()
            t_factor, t_total = e0.elapsed_time(e1) / 1e3, e0.elapsed_time(e2) / 1e3
        else:
            t_factor, t_total = t1 - t0, time.perf_counter() - t0
        tt = torch.tensor([t_factor, t_total], dtype=torch.float64, device=self.dev)
        if self.W > 1:
            self.comm.all_reduce(tt, tt, op="max")
        t_factor, t_total = float(tt[0]), float(tt[1])
        n = float(self.n)
        flops = 2.0 / 3.0 * n ** 3 + 1.5 * n ** 2
        self.x = x
        return {"n": self.n, "nb": self.nb, "world": self.W, "grid": [1, self.W], "passed": bool(history[-1] < tol),
                "scaled_residual": history[-1], "residual_history": history, "refinement_iterations": len(history) - 1,
                "seconds": t_total, "factor_seconds": t_factor, "gflops": flops / t_total / 1e9,
                "factor_gflops": (2.0 / 3.0 * n ** 3) / max(t_factor, 1e-12) / 1e9,
                "own_kernel_launches": self._launches() - self.launches0,
                "precision": "bf16 operands / fp32 accumulate LU, fp64 iterative refinement",
                "timing": "cuda events, max over ranks" if self.cuda else "wall clock (CPU)"}


def heap_bytes_for(n: int, nb: int) -> int:
    """Symmetric heap needed by :class:`HPLMxP`: the 32 MB control region, the staging carve-out (min(128 MB, half of the rest), see
    native/coll/comm.cpp) and the user part: panel (bf16) + right-hand side (fp32) + slack."""
    need = n * nb * 2 + n * 4 + (16 << 20)
    return (32 << 20) + max(2 * need, need + (128 << 20))


def run(comm, n: int, nb: int = 2048, seed: int = 42, max_refine: int = 50) -> dict:
    h = HPLMxP(comm, n, nb, seed)
    out = h.solve(max_refine=max_refine)
    if not out["passed"]:
        raise HPLError(f"HPL-MxP did not reach the scaled residual bound: {out['residual_history']}")
    return out


def reference_solution_error(h: HPLMxP, x: Optional[torch.Tensor] = None) -> float:
    """Test helper (small n): relative error of x against a dense fp64 solve of the regenerated matrix, on rank-local data only for W = 1."""
    assert h.W == 1
    a = torch.cat([h._a_block(j) for j in range(h.nblk)], dim=1)
    xr = torch.linalg.solve(a, h.b)
    x = h.x if x is None else x
    return float((x - xr).abs().max() / xr.abs().max())
