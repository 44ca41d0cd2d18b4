"""Global-resource pre-load ("cascade"): make every image/artefact available on the box.

Behavioural parity with /root/reference/cascade/cascade.py (bounded concurrent
pulls gated by leases on ``<sha1(resource)>.<i>`` for i < concurrency :574-646,
lease renewal :227-252, retry on transient registry errors with jittered
exponential backoff capped at 300 s :98-105,527-546, perf events
``cascade:start / pull-start / pull-end / gr-done`` :197-224,524-571,659) and
``scripts/shipyard_cascade.sh`` + ``wait_for_images.sh`` (block until loaded).

Re-designed for one box: a "pull" resolves the image to a local artefact (an
exported image tarball / directory in the image store, or a real ``docker pull`` /
``singularity pull`` when those runtimes exist) and streams its bytes through
the native staging arena (``libshipyard_stage``: pinned chunks -> cudaMemcpyAsync
on copy streams) so the artefact is page-cache/HBM resident before the first
task step, with at most ``concurrent_source_downloads`` transfers in flight.
"""
from __future__ import annotations

import hashlib
import os
import random
import re
import shutil
import subprocess
import threading
import time
from typing import Callable, Optional

from ..state.store import Store

TRANSIENT_ERRORS = ("toomanyrequests", "connection reset by peer", "error pulling image configuration",
                    "error parsing http 404 response body", "received unexpected http status",
                    "tls handshake timeout", "temporarily unavailable", "i/o timeout")
MAX_PULL_RETRIES = 25
LEASE_S, RENEW_S = 60.0, 15.0


def image_store_dir(state_dir: str) -> str:
    return os.environ.get("SHIPYARD_IMAGE_DIR") or os.path.join(state_dir, "images")


def artefact_name(resource: str) -> str:
    kind, _, name = resource.partition(":")
    return kind + "-" + re.sub(r"[^a-zA-Z0-9_.-]", "_", name)


def find_artefact(state_dir: str, resource: str) -> Optional[str]:
    base = os.path.join(image_store_dir(state_dir), artefact_name(resource))
    for cand in (base, base + ".tar", base + ".sif", base + ".tar.gz"):
        if os.path.exists(cand):
            return cand
    return None


def is_transient(msg: str) -> bool:
    m = msg.lower()
    return any(e in m for e in TRANSIENT_ERRORS)


def backoff_delays(base: float = 1.0, cap: float = 300.0, rng: Optional[random.Random] = None):
    """Jittered exponential backoff: uniform(0, min(cap, base*2^n))."""
    rng = rng or random.Random()
    n = 0
    while True:
        yield rng.uniform(0, min(cap, base * (2 ** n)))
        n += 1


class Cascade:
    def __init__(self, store: Store, pool_id: str, node_id: str = "box", concurrency: int = 10,
                 device: Optional[int] = None, puller: Optional[Callable] = None, sleep=time.sleep,
                 strict: Optional[bool] = None, fallback_registry: Optional[str] = None):
        self.store, self.pool_id, self.node_id = store, pool_id, node_id
        self.concurrency = max(1, int(concurrency))
        self.device, self.sleep = device, sleep
        self.puller = puller or self._default_pull
        self.strict = bool(int(os.environ.get("SHIPYARD_STRICT_IMAGES", "0"))) if strict is None else strict
        self.fallback_registry = fallback_registry
        self.holder = f"cascade-{os.getpid()}-{node_id}"
        self.errors: dict = {}
        self._stager = None
        self._lock = threading.Lock()

    # -- registry of resources to load (the reference's global-resource table) -------------
    @staticmethod
    def populate(store: Store, pool_id: str, resources: list[str]) -> None:
        store.delete("globalresource", pool_id)
        for r in resources:
            store.insert("globalresource", pool_id, hashlib.sha1(r.encode()).hexdigest(),
                         {"resource": r, "state": "pending", "size": None, "seconds": None, "error": None}, replace=True)

    def resources(self) -> list[dict]:
        return self.store.query("globalresource", self.pool_id)

    def _event(self, event: str, msg: Optional[str] = None) -> None:
        self.store.record_event("cascade", event, pool=self.pool_id, node=self.node_id, message=msg)

    # -- pulling --------------------------------------------------------------------------------
    def _stage_bytes(self, path: str) -> int:
        """Read the artefact through the native stager's pinned arena with bounded concurrency (page cache warm + integrity of every
        file proven by a full read; HBM when a device is set).  A stager failure is a pull failure: it surfaces in the resource's
        error column and counts against the retry budget — it is never replaced by a silent os.path.getsize()."""
        from ..ops.stage import Stager
        files = [path] if os.path.isfile(path) else [os.path.join(d, f) for d, _, fs in os.walk(path) for f in fs]
        total, t0 = 0, time.time()
        with self._lock:
            if self._stager is None:
                self._stager = Stager(self.device, arena_bytes=64 << 20, concurrency=min(4, self.concurrency))
        tickets = [(self._stager.submit_file(f), f) for f in files if os.path.getsize(f) > 0]
        for t, f in tickets:
            try:
                self._stager.wait(t)
                total += self._stager.query(t).bytes
            finally:
                self._stager.release(t)
        dt = max(time.time() - t0, 1e-9)
        self._event("stage", f"path={path},bytes={total},seconds={dt:.4f},mb_per_s={total / dt / 1e6:.1f},device={self.device}")
        return total

    def _default_pull(self, resource: str) -> int:
        """Resolve + load one resource; returns its size in bytes; raises RuntimeError(msg) on failure."""
        kind, _, name = resource.partition(":")
        art = find_artefact(self.store.root, resource)
        if art:
            return self._stage_bytes(art)
        exe = shutil.which("docker" if kind == "docker" else "singularity")
        if exe:
            cmd = [exe, "pull", name] if kind == "docker" else [exe, "pull", "--force", name]
            p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if p.returncode != 0:
                raise RuntimeError(p.stdout[-2000:])
            return 0
        if self.strict:
            raise RuntimeError(f"no artefact for {resource} in {image_store_dir(self.store.root)} and no {kind} runtime on the box")
        return 0   # virtual image: the process sandbox runs on the host filesystem

    def _acquire_slot(self, resource: str) -> Optional[str]:
        h = hashlib.sha1(resource.encode()).hexdigest()
        order = list(range(self.concurrency))
        random.shuffle(order)
        for i in order:
            name = f"cascade-{self.pool_id}-{h}.{i}"
            if self.store.acquire_lease(name, self.holder, LEASE_S):
                return name
        return None

    def _pull_one(self, ent: dict) -> bool:
        resource = ent["resource"]
        lease = None
        while lease is None:
            lease = self._acquire_slot(resource)
            if lease is None:
                self.sleep(0.05)
        stop = threading.Event()

        def renew():
            while not stop.wait(RENEW_S):
                self.store.renew_lease(lease, self.holder, LEASE_S)
        th = threading.Thread(target=renew, daemon=True)
        th.start()
        try:
            self._event("pull-start", resource)
            t0 = time.time()
            delays = backoff_delays()
            tried_fallback = False
            for attempt in range(MAX_PULL_RETRIES):
                try:
                    size = self.puller(resource)
                    dt = time.time() - t0
                    self.store.merge("globalresource", self.pool_id, ent["_rk"], {"state": "loaded", "size": size, "seconds": dt, "error": None})
                    self._event("pull-end", f"{resource},size={size},seconds={dt:.3f}")
                    return True
                except RuntimeError as e:
                    msg = str(e)
                    if is_transient(msg) and attempt + 1 < MAX_PULL_RETRIES:
                        self.sleep(next(delays))
                        continue
                    if self.fallback_registry and not tried_fallback and resource.startswith("docker:"):
                        tried_fallback = True     # retry the same image through the mirror registry
                        try:
                            size = self.puller(f"docker:{self.fallback_registry}/{resource.partition(':')[2]}")
                            self.store.merge("globalresource", self.pool_id, ent["_rk"], {"state": "loaded", "size": size,
                                             "seconds": time.time() - t0, "error": None, "via": "fallback_registry"})
                            self._event("pull-end", f"{resource},fallback=1")
                            return True
                        except RuntimeError as e2:
                            msg = str(e2)
                    self.errors[resource] = msg
                    self.store.merge("globalresource", self.pool_id, ent["_rk"], {"state": "failed", "error": msg[-500:]})
                    return False
            return False
        finally:
            stop.set()
            self.store.release_lease(lease, self.holder)

    def run(self, block: bool = True) -> bool:
        """Load every pending resource with bounded concurrency.  Returns True when all loaded."""
        self._event("start")
        pending = [e for e in self.resources() if e["state"] != "loaded"]
        results: list = []

        def work(chunk):
            for e in chunk:
                results.append(self._pull_one(e))
        n = max(1, min(self.concurrency, len(pending)))
        threads = [threading.Thread(target=work, args=(pending[i::n],), daemon=True) for i in range(n)]
        for t in threads:
            t.start()
        if not block:
            return True
        for t in threads:
            t.join()
        ok = all(results) if results else True
        if ok:
            self._event("gr-done", f"nglobalresources={len(self.resources())}")
        if self._stager is not None:
            self._stager.close(); self._stager = None
        return ok

    def all_loaded(self) -> bool:
        return all(e["state"] == "loaded" for e in self.resources())


def wait_for_images(store: Store, pool_id: str, timeout: float = 600.0, poll: float = 0.1) -> bool:
    """Block until every global resource of the pool is loaded (job-prep gate, wait_for_images.sh)."""
    t0 = time.time()
    while time.time() - t0 < timeout:
        ents = store.query("globalresource", pool_id)
        if all(e["state"] == "loaded" for e in ents):
            return True
        if any(e["state"] == "failed" for e in ents):
            return False
        time.sleep(poll)
    return False
