"""ctypes binding for ``libshipyard_stage`` (pinned arena + async H2D mover).

``Stager(device=None)`` is the host-only mode used on CPU boxes; with a device index the
arena is cudaHostAlloc'd and copies run on per-worker CUDA copy streams.  Used by the
image/artefact pre-loader (cascade equivalent: /root/reference/cascade/cascade.py:500-646 pulls images with
bounded concurrency), ``shipyard data ingress`` (/root/reference/convoy/data.py:492-876) and the recipes' input pipelines.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

_LIB = None


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_native", "libshipyard_stage.so")


def load() -> C.CDLL:
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            from .._build import ensure_built
            ensure_built(["stage"])
        lib = C.CDLL(p)
        vp, sz, i, lg, db = C.c_void_p, C.c_size_t, C.c_int, C.c_long, C.c_double
        lib.sy_stage_last_error.restype = C.c_char_p
        lib.sy_stage_create.argtypes = [C.POINTER(vp), i, sz, i, i]
        lib.sy_stage_destroy.argtypes = [vp]
        lib.sy_stage_submit_file.argtypes = [vp, C.c_char_p, vp, sz, sz]; lib.sy_stage_submit_file.restype = lg
        lib.sy_stage_submit_host.argtypes = [vp, vp, sz, vp]; lib.sy_stage_submit_host.restype = lg
        lib.sy_stage_submit_copy.argtypes = [vp, C.c_char_p, C.c_char_p, sz, sz]; lib.sy_stage_submit_copy.restype = lg
        lib.sy_stage_submit_pinned.argtypes = [vp, vp, sz, vp, vp]; lib.sy_stage_submit_pinned.restype = lg
        lib.sy_stage_pinned_alloc.argtypes = [vp, sz]; lib.sy_stage_pinned_alloc.restype = vp
        lib.sy_stage_pinned_free.argtypes = [vp, vp]
        lib.sy_stage_wait.argtypes = [vp, lg, db]
        lib.sy_stage_stream_wait.argtypes = [vp, lg, vp]
        lib.sy_stage_ptr.argtypes = [vp, lg]; lib.sy_stage_ptr.restype = vp
        lib.sy_stage_query.argtypes = [vp, lg, C.POINTER(C.c_ulonglong), C.POINTER(db)]
        lib.sy_stage_release.argtypes = [vp, lg]
        lib.sy_stage_stats.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(db)]
        _LIB = lib
    return _LIB


class StageError(RuntimeError):
    pass


@dataclass
class TicketInfo:
    state: str
    bytes: int
    done_bytes: int
    queue_seconds: float
    transfer_seconds: float


_STATES = {0: "queued", 1: "running", 2: "done", 3: "failed"}


class Stager:
    def __init__(self, device: Optional[int] = None, arena_bytes: int = 256 << 20, concurrency: int = 4,
                 chunks_per_worker: int = 2):
        self.lib = load()
        self.device = -1 if device is None else int(device)
        h = C.c_void_p()
        rc = self.lib.sy_stage_create(C.byref(h), self.device, int(arena_bytes), int(concurrency), int(chunks_per_worker))
        if rc != 0:
            raise StageError(f"stage create failed ({rc}): {self.lib.sy_stage_last_error().decode()}")
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.sy_stage_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def submit_file(self, path: str, dptr: int = 0, offset: int = 0, nbytes: int = 0) -> int:
        t = self.lib.sy_stage_submit_file(self._h, path.encode(), C.c_void_p(dptr), offset, nbytes)
        if t < 0:
            raise StageError(self.lib.sy_stage_last_error().decode())
        return int(t)

    def submit_copy(self, src: str, dst: str, offset: int = 0, nbytes: int = 0) -> int:
        """File -> file copy on a worker thread (copy_file_range, bounce through the arena where the kernel refuses): the task-side data
        mover's primitive.  Mode and mtime are carried over like shutil.copy2."""
        t = self.lib.sy_stage_submit_copy(self._h, os.fsencode(src), os.fsencode(dst), offset, nbytes)
        if t < 0:
            raise StageError(self.lib.sy_stage_last_error().decode())
        return int(t)

    def submit_host(self, host_ptr: int, nbytes: int, dptr: int = 0) -> int:
        t = self.lib.sy_stage_submit_host(self._h, C.c_void_p(host_ptr), nbytes, C.c_void_p(dptr))
        if t < 0:
            raise StageError(self.lib.sy_stage_last_error().decode())
        return int(t)

    def submit_pinned(self, host_ptr: int, nbytes: int, dptr: int, wait_event: int = 0) -> int:
        """Page-locked source -> device with ONE cudaMemcpyAsync on a worker copy stream (no bounce through the arena); the copy
        stream first waits for `wait_event` (a raw cudaEvent_t, e.g. ``torch.cuda.Event.cuda_event``) when given."""
        t = self.lib.sy_stage_submit_pinned(self._h, C.c_void_p(host_ptr), nbytes, C.c_void_p(dptr), C.c_void_p(wait_event or 0))
        if t < 0:
            raise StageError(self.lib.sy_stage_last_error().decode())
        return int(t)

    def wait(self, ticket: int, timeout: float = -1.0) -> bool:
        rc = self.lib.sy_stage_wait(self._h, ticket, float(timeout))
        if rc == 1:
            return False
        if rc != 0:
            raise StageError(f"ticket {ticket}: {self.lib.sy_stage_last_error().decode()}")
        return True

    def stream_wait(self, ticket: int, cuda_stream: int) -> None:
        rc = self.lib.sy_stage_stream_wait(self._h, ticket, C.c_void_p(cuda_stream))
        if rc != 0:
            raise StageError(f"ticket {ticket}: {self.lib.sy_stage_last_error().decode()}")

    def ptr(self, ticket: int) -> int:
        return int(self.lib.sy_stage_ptr(self._h, ticket) or 0)

    def query(self, ticket: int) -> TicketInfo:
        out = (C.c_ulonglong * 3)(); secs = (C.c_double * 2)()
        if self.lib.sy_stage_query(self._h, ticket, out, secs) != 0:
            raise StageError(f"unknown ticket {ticket}")
        return TicketInfo(_STATES[int(out[0])], int(out[1]), int(out[2]), float(secs[0]), float(secs[1]))

    def read_host(self, ticket: int) -> bytes:
        """Host-only mode: the staged bytes."""
        assert self.device < 0
        info = self.query(ticket)
        return C.string_at(self.ptr(ticket), info.bytes)

    def release(self, ticket: int) -> None:
        self.lib.sy_stage_release(self._h, ticket)

    def stats(self) -> dict:
        out = (C.c_ulonglong * 4)(); busy = C.c_double()
        self.lib.sy_stage_stats(self._h, out, C.byref(busy))
        return {"bytes_staged": int(out[0]), "memcpy_calls": int(out[1]), "arena_bytes": int(out[2]),
                "chunk_bytes": int(out[3]), "busy_seconds": float(busy.value)}


class StagedInputs:
    """Task input files on their way to (or in) HBM: ``tensors[path]`` is a uint8 CUDA tensor over the stager-owned buffer, valid
    for kernels enqueued after :meth:`chain` on that stream (or after :meth:`wait` on the host)."""

    def __init__(self, stager: Stager, entries: list):
        self.stager, self.entries = stager, entries          # entries: (path, ticket, nbytes)
        self.t_submit = __import__("time").time()

    def chain(self, cuda_stream: int) -> None:
        """Event-chain `cuda_stream` behind every copy (sy_stage_stream_wait): the host does not block on the transfers."""
        for _, t, _ in self.entries:
            self.stager.stream_wait(t, cuda_stream)

    def wait(self) -> None:
        for _, t, _ in self.entries:
            self.stager.wait(t)

    def done(self) -> bool:
        return all(self.stager.query(t).state in ("done", "failed") for _, t, _ in self.entries)

    @property
    def nbytes(self) -> int:
        return sum(n for _, _, n in self.entries)

    def tensors(self) -> dict:
        import torch
        out = {}
        for path, t, n in self.entries:
            view = type("_V", (), {})()
            view.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (self.stager.ptr(t), False), "version": 3, "strides": None}
            view._owner = self
            out[path] = torch.as_tensor(view, device=f"cuda:{self.stager.device}")
        return out

    def summary(self) -> dict:
        secs = max((self.stager.query(t).transfer_seconds for _, t, _ in self.entries), default=0.0)
        return {"files": len(self.entries), "bytes": self.nbytes, "slowest_ticket_s": round(secs, 4)}


def stage_task_inputs(device: int, manifest: Optional[str] = None, stager: Optional[Stager] = None, concurrency: int = 4) -> Optional[StagedInputs]:
    """Push every file of the task's ``input_data`` (the manifest the ingress prologue wrote, $SHIPYARD_STAGE_MANIFEST) through
    file -> pinned arena -> HBM tickets and return immediately; the caller runs its first step(s) and chains the consumer stream
    behind the tickets when it needs the data.  Reference counterpart: blobxfer downloads in the task prologue
    (/root/reference/convoy/data.py:219-291, scripts/shipyard_blobxfer.sh) — data only ever reached the node's disk."""
    import json
    manifest = manifest or os.environ.get("SHIPYARD_STAGE_MANIFEST")
    if not manifest or not os.path.exists(manifest):
        return None
    with open(manifest) as f:
        files = [e for e in json.load(f).get("files", []) if e.get("bytes", 0) > 0 and os.path.exists(e["path"])]
    if not files:
        return None
    st = stager or Stager(device, arena_bytes=concurrency * 2 * (16 << 20), concurrency=concurrency)
    return StagedInputs(st, [(e["path"], st.submit_file(e["path"]), int(e["bytes"])) for e in files])
