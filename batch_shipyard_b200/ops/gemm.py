"""ctypes binding + autograd wrappers for ``libshipyard_gemm`` (tcgen05/TMEM/TMA bf16 GEMM).

``gemm_tn(a, b)`` computes ``a @ b.T`` for row-major bf16 ``a[M,K]``, ``b[N,K]`` — the shape of a
``Linear`` and of a 1x1 convolution on NHWC activations (``X[M=N*H*W, Cin] @ W[Cout, Cin]^T``).
With ``stats=`` the kernel's epilogue also accumulates the per-channel sum and sum of squares
of the output, i.e. the train-mode BatchNorm statistics, saving one full pass over the activations.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_LIB = None


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_native", "libshipyard_gemm.so")


def load() -> C.CDLL:
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            from .._build import ensure_built
            ensure_built(["gemm"])
        if not os.path.exists(p):
            raise RuntimeError(f"{p} missing: run `python native/build.py gemm` (no fallback on GPU)")
        lib = C.CDLL(p)
        lib.sy_gemm_last_error.restype = C.c_char_p
        lib.sy_gemm_launch_count.restype = C.c_ulonglong
        lib.sy_gemm_bf16_tn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.sy_gemm_bf16_tn_allreduce.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _LIB = lib
        try:
            from ..parallel.ddp import register_launch_counter
            register_launch_counter(lambda: int(lib.sy_gemm_launch_count()))
        except Exception:  # noqa: BLE001
            pass
    return _LIB


def launch_count() -> int:
    return int(load().sy_gemm_launch_count())


def _rows_ok(t: torch.Tensor) -> bool:
    return t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
            stats: Optional[torch.Tensor] = None, block_n: int = 0, max_ctas: int = 0) -> torch.Tensor:
    """out[M,N] = a[M,K] @ b[N,K]^T (+ bias[N]); bf16, fp32 accumulation in TMEM.
    stats: zero-initialised float32[2*N] receiving column sums / sums of squares of `out`."""
    assert a.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.shape[1] == b.shape[1]
    m, k = a.shape
    n = b.shape[0]
    if k % 8:                                   # TMA needs 16-byte row pitches: zero-pad K (zeros do not change the product)
        pad = 8 - k % 8
        a = torch.nn.functional.pad(a, (0, pad)); b = torch.nn.functional.pad(b, (0, pad)); k += pad
    if not _rows_ok(a):
        a = a.contiguous()
    if not _rows_ok(b):
        b = b.contiguous()
    if out is None:
        ldc = (n + 7) // 8 * 8
        buf = torch.empty((m, ldc), dtype=torch.bfloat16, device=a.device)
        out = buf[:, :n] if ldc != n else buf
    assert out.shape == (m, n) and out.dtype == torch.bfloat16 and _rows_ok(out)
    if stats is not None:
        assert stats.dtype == torch.float32 and stats.numel() == 2 * n and stats.is_contiguous()
    if bias is not None:
        assert bias.dtype == torch.bfloat16 and bias.numel() == n and bias.is_contiguous()
    lib = load()
    rc = lib.sy_gemm_bf16_tn(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), m, n, k,
                             a.stride(0), b.stride(0), out.stride(0), C.c_void_p(bias.data_ptr() if bias is not None else 0),
                             C.c_void_p(stats.data_ptr() if stats is not None else 0), block_n, max_ctas,
                             C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"sy_gemm_bf16_tn failed ({rc}): {lib.sy_gemm_last_error().decode()}")
    return out


class _LinearTN(torch.autograd.Function):
    """y = x @ w^T (+ b): forward and dgrad on the tcgen05 kernel; wgrad (both operands MN-major) via cuBLAS."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return gemm_tn(x, w, bias=b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = gemm_tn(dy, w.t().contiguous()) if ctx.needs_input_grad[0] else None     # dX = dY @ W  (W^T is small)
        dw = (dy.t() @ x) if ctx.needs_input_grad[1] else None
        db = dy.float().sum(0).to(dy.dtype) if ctx.has_bias else None
        return dx, dw, db


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _LinearTN.apply(x, w, b)


class _Conv1x1NHWC(torch.autograd.Function):
    """1x1 stride-1 convolution on channels_last bf16 activations as one GEMM; the epilogue produces the
    BatchNorm statistics of the output (returned as a non-differentiable float32[2*Cout])."""

    @staticmethod
    def forward(ctx, x, w, want_stats):
        n, cin, h, wd = x.shape
        cout = w.shape[0]
        x2 = x.permute(0, 2, 3, 1).reshape(n * h * wd, cin)            # view: NHWC storage
        w2 = w.permute(0, 2, 3, 1).reshape(cout, cin)
        stats = torch.zeros(2 * cout, dtype=torch.float32, device=x.device) if want_stats else None
        y2 = gemm_tn(x2, w2, stats=stats)
        ctx.save_for_backward(x, w)
        y = y2.view(n, h, wd, cout).permute(0, 3, 1, 2)
        ctx.mark_non_differentiable(stats) if stats is not None else None
        return (y, stats) if want_stats else (y, None)

    @staticmethod
    def backward(ctx, dy, _dstats):
        x, w = ctx.saved_tensors
        n, cin, h, wd = x.shape
        cout = w.shape[0]
        dy2 = dy.permute(0, 2, 3, 1).reshape(n * h * wd, cout)
        if not _rows_ok(dy2):
            dy2 = dy2.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wt = w.permute(0, 2, 3, 1).reshape(cout, cin).t().contiguous()          # [Cin, Cout], small
            dx = gemm_tn(dy2, wt).view(n, h, wd, cin).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            x2 = x.permute(0, 2, 3, 1).reshape(n * h * wd, cin)
            dw = (dy2.t() @ x2).view(cout, 1, 1, cin).permute(0, 3, 1, 2)
        return dx, dw, None


def conv1x1_nhwc(x: torch.Tensor, w: torch.Tensor, want_stats: bool = False):
    return _Conv1x1NHWC.apply(x, w, want_stats)


def gemm_tn_allreduce(comm, a: torch.Tensor, b: torch.Tensor, out_sym: torch.Tensor, block_n: int = 0) -> torch.Tensor:
    """K10: ``out_sym += a @ b.T`` summed over all ranks, ONE kernel (tcgen05 GEMM whose epilogue issues
    multimem.red into the NVLS multicast mapping of ``out_sym``; ends with a cross-GPU flag barrier).
    ``out_sym``: fp32 [M, N] from ``comm.alloc`` and zero on every rank before the call; a/b: this rank's K-shard."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and out_sym.dtype == torch.float32
    m, k = a.shape
    n = b.shape[0]
    assert tuple(out_sym.shape) == (m, n) and out_sym.is_contiguous() and _rows_ok(a) and _rows_ok(b)
    lib = load()
    from . import coll as _coll
    clib = _coll.load()
    clib.sy_comm_device_view.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    clib.sy_comm_device_view.restype = C.c_size_t
    size = clib.sy_comm_device_view(comm._h, None, 0)
    view = (C.c_uint8 * size)()
    clib.sy_comm_device_view(comm._h, view, size)
    rc = lib.sy_gemm_bf16_tn_allreduce(view, size, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), comm.heap_offset(out_sym), m, n, k,
                                       a.stride(0), b.stride(0), out_sym.stride(0), block_n,
                                       C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"sy_gemm_bf16_tn_allreduce failed ({rc}): {lib.sy_gemm_last_error().decode()}")
    return out_sym
