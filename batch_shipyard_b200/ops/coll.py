"""Python binding for ``libshipyard_coll`` (ctypes, C ABI).

``Communicator`` owns one symmetric heap per rank and exposes the fused
collective kernels on torch tensors.  On a box without GPUs (or with
``device=None``) the same API runs on the host shared-memory stub transport,
which is how the CPU test-suite exercises world_size>1.

Parity: the reference has no communication code of its own; it generates an
``mpirun`` line and lets the container's MPI/NCCL do the work
(/root/reference/convoy/batch.py:4362-4486).  This module is the data plane
that replaces those third-party libraries on one NVSwitch box.
"""
from __future__ import annotations

import ctypes as C
import os
import uuid
from typing import Optional

import torch

_LIB = None

F32, BF16, F16, F64, I32, I64, U8 = range(7)
SUM, MAX, MIN, PROD = range(4)
ALGO_AUTO, ALGO_LL, ALGO_ONESHOT, ALGO_TWOSHOT_P2P, ALGO_TWOSHOT_NVLS = range(5)
TRANSPORT_AUTO, TRANSPORT_STUB, TRANSPORT_P2P, TRANSPORT_NVLS = range(4)
TRANSPORT_NAMES = {1: "stub", 2: "p2p", 3: "nvls"}

_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16, torch.float64: F64,
       torch.int32: I32, torch.int64: I64, torch.uint8: U8}
_OPS = {"sum": SUM, "max": MAX, "min": MIN, "prod": PROD}
_ALGOS = {"auto": ALGO_AUTO, "ll": ALGO_LL, "oneshot": ALGO_ONESHOT, "twoshot_p2p": ALGO_TWOSHOT_P2P,
          "twoshot_nvls": ALGO_TWOSHOT_NVLS}


class HaloDesc(C.Structure):
    _fields_ = [("peer", C.c_int), ("sig_idx", C.c_int), ("dst_off", C.c_long),
                ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("sx", C.c_long), ("sy", C.c_long), ("sz", C.c_long), ("src_elem_off", C.c_long)]


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_native",
                        "libshipyard_coll.so")


def load() -> C.CDLL:
    """Load the native library; build it on demand when the sources are present."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        from .._build import ensure_built
        ensure_built(["coll"])
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: run `python native/build.py coll` (the collectives have no "
                           "Python fallback by design)")
    lib = C.CDLL(p, mode=C.RTLD_LOCAL)
    vp, sz, i, f = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    lib.sy_last_error.restype = C.c_char_p
    lib.sy_comm_init.argtypes = [C.POINTER(vp), i, i, C.c_char_p, i, sz, i]
    lib.sy_comm_destroy.argtypes = [vp]
    for n in ("sy_comm_rank", "sy_comm_world", "sy_comm_transport", "sy_comm_has_multicast", "sy_comm_status"):
        getattr(lib, n).argtypes = [vp]
    lib.sy_comm_launch_count.argtypes = [vp]; lib.sy_comm_launch_count.restype = C.c_uint64
    lib.sy_sym_alloc.argtypes = [vp, sz]; lib.sy_sym_alloc.restype = vp
    lib.sy_sym_reset.argtypes = [vp]
    lib.sy_heap_base.argtypes = [vp, i]; lib.sy_heap_base.restype = vp
    lib.sy_mc_base.argtypes = [vp]; lib.sy_mc_base.restype = vp
    lib.sy_heap_bytes.argtypes = [vp]; lib.sy_heap_bytes.restype = sz
    lib.sy_is_symmetric.argtypes = [vp, vp]
    lib.sy_set_tuning.argtypes = [vp, C.c_char_p, C.c_long]
    lib.sy_get_tuning.argtypes = [vp, C.c_char_p]; lib.sy_get_tuning.restype = C.c_long
    lib.sy_allreduce.argtypes = [vp, vp, vp, sz, i, i, f, i, i, vp]
    lib.sy_reduce_scatter.argtypes = [vp, vp, vp, sz, i, i, f, i, vp]
    lib.sy_allgather.argtypes = [vp, vp, vp, sz, i, vp]
    lib.sy_broadcast.argtypes = [vp, vp, vp, sz, i, i, vp]
    lib.sy_alltoall.argtypes = [vp, vp, vp, sz, i, vp]
    lib.sy_reduce.argtypes = [vp, vp, vp, sz, i, i, i, vp]
    lib.sy_gather.argtypes = [vp, vp, vp, sz, i, i, vp]
    lib.sy_scatter.argtypes = [vp, vp, vp, sz, i, i, vp]
    lib.sy_barrier.argtypes = [vp, vp]
    lib.sy_put_signal.argtypes = [vp, vp, sz, sz, i, i, vp]
    lib.sy_wait_signal.argtypes = [vp, i, C.c_uint32, vp]
    lib.sy_halo_exchange.argtypes = [vp, vp, i, C.POINTER(HaloDesc), i, C.POINTER(C.c_int), i, vp]
    lib.sy_shard_begin.argtypes = [vp, sz, i]; lib.sy_shard_begin.restype = sz
    lib.sy_shard_count.argtypes = [vp, sz, i]; lib.sy_shard_count.restype = sz
    lib.sy_fused_allreduce_sgd.argtypes = [vp, vp, i, vp, i, vp, vp, vp, sz, i, vp]
    lib.sy_allreduce_fp8_blockscaled.argtypes = [vp, vp, i, vp, vp, sz, f, vp]
    _LIB = lib
    return lib


class CollError(RuntimeError):
    pass


class _CudaView:
    """Minimal __cuda_array_interface__ carrier so torch can wrap heap memory."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3, "strides": None}
        self._owner = owner


def default_session(tag: str = "") -> str:
    """A session id every rank of a torchrun/shipyard launch derives identically."""
    env = os.environ
    base = env.get("SHIPYARD_COLL_SESSION") or env.get("TORCHELASTIC_RUN_ID") or ""
    port = env.get("MASTER_PORT", "0")
    return f"{base}-{port}-{tag}" if (base or port != "0") else f"solo-{os.getpid()}-{tag}"


class Communicator:
    """One rank's endpoint of a single-box communicator."""

    def __init__(self, rank: int = 0, world: int = 1, session: Optional[str] = None,
                 device: Optional[int] = None, heap_bytes: int = 0, transport: str = "auto"):
        self.lib = load()
        self.rank, self.world = int(rank), int(world)
        if session is None:
            session = default_session() if world > 1 else f"solo-{os.getpid()}-{uuid.uuid4().hex[:8]}"
        self.session = session
        self.device_index = -1 if device is None else int(device)
        tr = {"auto": TRANSPORT_AUTO, "stub": TRANSPORT_STUB, "p2p": TRANSPORT_P2P, "nvls": TRANSPORT_NVLS}[transport]
        h = C.c_void_p()
        rc = self.lib.sy_comm_init(C.byref(h), self.rank, self.world, session.encode(), self.device_index,
                                   int(heap_bytes), tr)
        if rc != 0:
            raise CollError(f"sy_comm_init failed ({rc}): {self.lib.sy_last_error().decode()}")
        self._h = h
        self.transport = TRANSPORT_NAMES[self.lib.sy_comm_transport(h)]
        self.is_stub = self.transport == "stub"
        self.has_multicast = bool(self.lib.sy_comm_has_multicast(h))
        self.heap_bytes = self.lib.sy_heap_bytes(h)
        self.torch_device = torch.device("cpu") if self.is_stub else torch.device("cuda", self.device_index)
        self._keep = []
        self._trace = None
        if os.environ.get("SHIPYARD_TRACE"):
            self._install_tracing(os.environ["SHIPYARD_TRACE"])

    # -- lifecycle ---------------------------------------------------------
    @classmethod
    def from_env(cls, device: Optional[int] = None, tag: str = "", **kw) -> "Communicator":
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if device is None and torch.cuda.is_available():
            device = int(os.environ.get("LOCAL_RANK", "0"))
        return cls(rank, world, default_session(tag), device, **kw)

    def close(self) -> None:
        if getattr(self, "_trace", None) is not None:
            try:
                self.flush_trace(final=True)
            except Exception:  # noqa: BLE001 - tracing must never break teardown
                pass
        if getattr(self, "_h", None) is not None:
            self.lib.sy_comm_destroy(self._h)
            self._h = None

    # -- tracing (JSONL trace per rank, device-side timings per collective, NVTX ranges) -------------
    _TRACED = ("all_reduce", "reduce_scatter", "all_gather", "broadcast", "all_to_all", "reduce", "gather", "scatter", "barrier",
               "halo_exchange", "fused_allreduce_sgd", "fused_allreduce_adam", "all_reduce_fp8")
    _BUCKETS_US = (5, 10, 20, 50, 100, 200, 500, 1000, 5000, 20000, 100000)

    def _install_tracing(self, path: str) -> None:
        """SHIPYARD_TRACE=<prefix>: every collective call is bracketed by CUDA events on its stream (time.perf_counter on the
        stub transport) and an NVTX range; records go to <prefix>.rank<r>.jsonl without ever synchronising the stream (events
        are harvested when they have completed, the rest at close()).  A latency histogram per operation is written next to
        the trace and, when SHIPYARD_STATE_DIR is set, under <state>/metrics/ for the Prometheus exporter."""
        self._trace = {"path": f"{path}.rank{self.rank}.jsonl", "pending": [], "hist": {}, "records": 0}
        for name in self._TRACED:
            fn = getattr(self, name, None)
            if fn is not None:
                setattr(self, name, self._traced(name, fn))

    def _traced(self, name, fn):
        import time as _time

        def wrapper(*a, **kw):
            tr = self._trace
            first = next((x for x in a if isinstance(x, torch.Tensor)), None)
            nbytes = 0 if first is None else first.numel() * first.element_size()
            cuda = self.torch_device.type == "cuda"
            if cuda and torch.cuda.is_current_stream_capturing():
                return fn(*a, **kw)                                  # inside a graph capture: nothing to time per call
            ts = _time.time()
            if cuda:
                stream = kw.get("stream") or torch.cuda.current_stream(self.torch_device)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.nvtx.range_push(f"shipyard:{name}:{nbytes}B")
                e0.record(stream)
                try:
                    return fn(*a, **kw)
                finally:
                    e1.record(stream)
                    torch.cuda.nvtx.range_pop()
                    tr["pending"].append((ts, name, nbytes, e0, e1))
                    if len(tr["pending"]) >= 256:
                        self.flush_trace()
            else:
                t0 = _time.perf_counter()
                try:
                    return fn(*a, **kw)
                finally:
                    self._trace_emit(ts, name, nbytes, (_time.perf_counter() - t0) * 1e6)
        wrapper.__name__ = name
        return wrapper

    def _trace_emit(self, ts: float, name: str, nbytes: int, us: float) -> None:
        import json as _json
        tr = self._trace
        with open(tr["path"], "a") as f:
            f.write(_json.dumps({"ts": round(ts, 6), "rank": self.rank, "world": self.world, "event": f"coll:{name}", "bytes": nbytes,
                                 "device_us": round(us, 2), "transport": self.transport}) + "\n")
        h = tr["hist"].setdefault(name, {"count": 0, "sum_us": 0.0, "buckets": [0] * (len(self._BUCKETS_US) + 1)})
        h["count"] += 1; h["sum_us"] += us
        h["buckets"][next((i for i, b in enumerate(self._BUCKETS_US) if us <= b), len(self._BUCKETS_US))] += 1
        tr["records"] += 1

    def flush_trace(self, final: bool = False) -> None:
        import json as _json
        tr = self._trace
        if tr is None:
            return
        keep = []
        for (ts, name, nbytes, e0, e1) in tr["pending"]:
            if final:
                e1.synchronize()
            if e1.query():
                self._trace_emit(ts, name, nbytes, e0.elapsed_time(e1) * 1e3)
            else:
                keep.append((ts, name, nbytes, e0, e1))
        tr["pending"] = keep
        if final:
            summary = {"rank": self.rank, "world": self.world, "transport": self.transport, "buckets_us": list(self._BUCKETS_US), "ops": tr["hist"]}
            with open(tr["path"].replace(".jsonl", ".hist.json"), "w") as f:
                _json.dump(summary, f)
            sd = os.environ.get("SHIPYARD_STATE_DIR")
            if sd:
                os.makedirs(os.path.join(sd, "metrics"), exist_ok=True)
                with open(os.path.join(sd, "metrics", f"coll-{os.getpid()}-rank{self.rank}.json"), "w") as f:
                    _json.dump(summary, f)

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise CollError(f"{what} failed ({rc}): {self.lib.sy_last_error().decode()}")

    def check_status(self) -> None:
        """Raise if the device-side watchdog recorded a flag-wait timeout."""
        st = self.lib.sy_comm_status(self._h)
        if st != 0:
            raise CollError(f"collective watchdog fired on rank {self.rank} (status {st}): a peer did not arrive")

    @property
    def launches(self) -> int:
        return int(self.lib.sy_comm_launch_count(self._h))

    def set_tuning(self, **kw) -> None:
        for k, v in kw.items():
            self._check(self.lib.sy_set_tuning(self._h, k.encode(), int(v)), f"set_tuning({k})")

    def get_tuning(self, k: str) -> int:
        return int(self.lib.sy_get_tuning(self._h, k.encode()))

    # -- symmetric memory --------------------------------------------------
    def alloc(self, shape, dtype=torch.float32) -> torch.Tensor:
        """Collective: allocate a tensor at the same heap offset on every rank."""
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        n = 1
        for s in shape:
            n *= int(s)
        esz = torch.empty((), dtype=dtype).element_size()
        nbytes = max(n * esz, 16)
        ptr = self.lib.sy_sym_alloc(self._h, nbytes)
        if not ptr:
            raise CollError(self.lib.sy_last_error().decode())
        if self.is_stub:
            buf = (C.c_uint8 * nbytes).from_address(ptr)
            t = torch.frombuffer(buf, dtype=torch.uint8, count=nbytes)
            self._keep.append(buf)
        else:
            t = torch.as_tensor(_CudaView(ptr, nbytes, self), device=self.torch_device)
        return t[: n * esz].view(dtype).view(shape)

    def reset_heap(self) -> None:
        self.lib.sy_sym_reset(self._h)

    def is_symmetric(self, t: torch.Tensor) -> bool:
        return bool(self.lib.sy_is_symmetric(self._h, C.c_void_p(t.data_ptr())))

    def heap_offset(self, t: torch.Tensor) -> int:
        base = self.lib.sy_heap_base(self._h, self.rank)
        off = t.data_ptr() - base
        if off < 0 or off >= self.heap_bytes:
            raise CollError("tensor is not in the symmetric heap")
        return off

    def shard_range(self, count: int, rank: Optional[int] = None) -> tuple[int, int]:
        r = self.rank if rank is None else rank
        return (int(self.lib.sy_shard_begin(self._h, count, r)), int(self.lib.sy_shard_count(self._h, count, r)))

    # -- helpers -----------------------------------------------------------
    def _stream(self, stream) -> C.c_void_p:
        if self.is_stub:
            return C.c_void_p(0)
        s = stream if stream is not None else torch.cuda.current_stream(self.torch_device)
        return C.c_void_p(s.cuda_stream)

    @staticmethod
    def _p(t: torch.Tensor) -> C.c_void_p:
        assert t.is_contiguous(), "collective operands must be contiguous"
        return C.c_void_p(t.data_ptr())

    # -- collectives -------------------------------------------------------
    def all_reduce(self, inp: torch.Tensor, out: Optional[torch.Tensor] = None, scale: float = 1.0,
                   op: str = "sum", algo: str = "auto", stream=None) -> torch.Tensor:
        out = inp if out is None else out
        assert out.numel() == inp.numel()
        self._check(self.lib.sy_allreduce(self._h, self._p(inp), self._p(out), inp.numel(), _DT[inp.dtype],
                                          _DT[out.dtype], float(scale), _OPS[op], _ALGOS[algo],
                                          self._stream(stream)), "all_reduce")
        return out

    def reduce_scatter(self, inp, out, scale: float = 1.0, op: str = "sum", stream=None):
        assert inp.numel() == out.numel() * self.world
        self._check(self.lib.sy_reduce_scatter(self._h, self._p(inp), self._p(out), out.numel(), _DT[inp.dtype],
                                               _DT[out.dtype], float(scale), _OPS[op], self._stream(stream)),
                    "reduce_scatter")
        return out

    def all_gather(self, inp, out, stream=None):
        assert out.numel() == inp.numel() * self.world and inp.dtype == out.dtype
        self._check(self.lib.sy_allgather(self._h, self._p(inp), self._p(out), inp.numel(), _DT[inp.dtype],
                                          self._stream(stream)), "all_gather")
        return out

    def broadcast(self, t, root: int = 0, stream=None):
        self._check(self.lib.sy_broadcast(self._h, self._p(t), self._p(t), t.numel(), _DT[t.dtype], root,
                                          self._stream(stream)), "broadcast")
        return t

    def all_to_all(self, inp, out, stream=None):
        assert inp.numel() == out.numel() and inp.numel() % self.world == 0
        self._check(self.lib.sy_alltoall(self._h, self._p(inp), self._p(out), inp.numel() // self.world,
                                         _DT[inp.dtype], self._stream(stream)), "all_to_all")
        return out

    def reduce(self, inp, out, root: int = 0, op: str = "sum", stream=None):
        self._check(self.lib.sy_reduce(self._h, self._p(inp), self._p(out), inp.numel(), _DT[inp.dtype], _OPS[op],
                                       root, self._stream(stream)), "reduce")
        return out

    def gather(self, inp, out, root: int = 0, stream=None):
        self._check(self.lib.sy_gather(self._h, self._p(inp), self._p(out), inp.numel(), _DT[inp.dtype], root,
                                       self._stream(stream)), "gather")
        return out

    def scatter(self, inp, out, root: int = 0, stream=None):
        self._check(self.lib.sy_scatter(self._h, self._p(inp), self._p(out), out.numel(), _DT[out.dtype], root,
                                        self._stream(stream)), "scatter")
        return out

    def barrier(self, stream=None) -> None:
        self._check(self.lib.sy_barrier(self._h, self._stream(stream)), "barrier")

    def put_signal(self, src: torch.Tensor, dst_off: int, peer: int, sig: int, stream=None) -> None:
        self._check(self.lib.sy_put_signal(self._h, self._p(src), dst_off, src.numel() * src.element_size(), peer,
                                           sig, self._stream(stream)), "put_signal")

    def wait_signal(self, sig: int, expected: int, stream=None) -> None:
        self._check(self.lib.sy_wait_signal(self._h, sig, expected, self._stream(stream)), "wait_signal")

    def halo_exchange(self, src: torch.Tensor, descs: list[HaloDesc], wait_sigs: list[int], stream=None) -> None:
        arr = (HaloDesc * max(1, len(descs)))(*descs)
        ws = (C.c_int * max(1, len(wait_sigs)))(*wait_sigs)
        self._check(self.lib.sy_halo_exchange(self._h, self._p(src), _DT[src.dtype], arr, len(descs), ws,
                                              len(wait_sigs), self._stream(stream)), "halo_exchange")

    def fused_allreduce_sgd(self, grads, params, master, mom, hyper, zero_grads: bool = True, stream=None) -> None:
        """grads/params: symmetric flat tensors; master/mom: local fp32 shards; hyper: 4 floats
        (lr, momentum, weight_decay, grad_scale) on the same device."""
        assert grads.numel() == params.numel() and grads.numel() % 8 == 0
        self._check(self.lib.sy_fused_allreduce_sgd(self._h, self._p(grads), _DT[grads.dtype], self._p(params),
                                                    _DT[params.dtype], self._p(master), self._p(mom), self._p(hyper),
                                                    grads.numel(), 1 if zero_grads else 0, self._stream(stream)),
                    "fused_allreduce_sgd")

    def fused_allreduce_adam(self, grad, param, m, v, hyper, scale: Optional[float] = None, zero_grad: bool = True, stream=None) -> None:
        """One launch: gradient all-reduce (x scale, default 1/world) + Adam on this rank's parameter replica.
        grad/param/m/v: fp32 flat tensors (numel % 4 == 0, <= 256 Ki); hyper: fp32 [lr, beta1, beta2, eps, step] on the device."""
        assert grad.dtype == param.dtype == m.dtype == v.dtype == torch.float32 and grad.numel() % 4 == 0
        self.lib.sy_fused_allreduce_adam.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                     C.c_float, C.c_int, C.c_void_p]
        self._check(self.lib.sy_fused_allreduce_adam(self._h, self._p(grad), self._p(param), self._p(m), self._p(v), self._p(hyper),
                                                     grad.numel(), float(1.0 / self.world if scale is None else scale),
                                                     1 if zero_grad else 0, self._stream(stream)), "fused_allreduce_adam")

    def all_reduce_fp8(self, inp, out_q, out_scales, scale: float = 1.0, stream=None) -> None:
        assert inp.numel() % 128 == 0 and out_q.numel() == inp.numel() and out_scales.numel() == inp.numel() // 32
        self._check(self.lib.sy_allreduce_fp8_blockscaled(self._h, self._p(inp), _DT[inp.dtype], self._p(out_q),
                                                          self._p(out_scales), inp.numel(), float(scale),
                                                          self._stream(stream)), "all_reduce_fp8")


def dequant_mx_fp8(q: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    """Reference dequantisation of (e4m3 bytes, e8m0 scale bytes per 32) -> fp32."""
    x = q.view(torch.float8_e4m3fn).to(torch.float32).view(-1, 32)
    s = torch.pow(2.0, scales.to(torch.float32) - 127.0).view(-1, 1)
    return (x * s).view(-1)
