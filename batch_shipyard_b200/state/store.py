"""Durable local state store: the stand-in for Azure Storage tables/queues/blobs/leases.

The reference keeps ALL state in the cloud (Batch objects + storage
tables/queues/blobs, /root/reference/convoy/storage.py:68-135, 424-488,
679-1745) so the CLI itself is stateless.  We keep that property: every
``shipyard`` invocation, the node agent, the federation daemon and the task
runner share one crash-safe SQLite database (WAL) plus a blob directory under
the state dir.  Same logical entities and concurrency primitives:

* tables   -> ``entities(kind, pk, rk)`` rows with an **etag** for compare-and-swap
* queues   -> visibility-timeout message queues (federation actions, slurm assignments)
* blobs    -> files (resource files, pickled/JSON job payloads, task output)
* leases   -> TTL'd named locks (image pre-load slots, federation leader election)
* perf     -> timing events (``nodeprep:start`` ... ``cascade:pull-end``, ``gr-done``)
"""
from __future__ import annotations

import json
import os
import shutil
import sqlite3
import threading
import time
import uuid
from contextlib import contextmanager
from typing import Any, Optional


class EntityExists(Exception):
    pass


class EtagMismatch(Exception):
    pass


class NotFound(KeyError):
    pass


def default_state_dir() -> str:
    return os.environ.get("SHIPYARD_STATE_DIR") or os.path.join(os.path.expanduser("~"), ".shipyard-b200")


def entity_names(prefix: str, account: str = "local", pool_id: Optional[str] = None) -> dict:
    """Names of the metadata containers/tables/queues for a storage_entity_prefix (3-63 chars each)."""
    sfx = f"-{account}-{pool_id.lower()}" if pool_id else ""
    names = {
        "blob_globalresources": f"{prefix}gr{sfx}", "blob_resourcefiles": f"{prefix}rf{sfx}",
        "blob_remotefs": f"{prefix}remotefs", "blob_monitoring": f"{prefix}monitor",
        "blob_federation": f"{prefix}fed", "blob_federation_global": f"{prefix}fedglobal",
        "table_images": f"{prefix}images", "table_globalresources": f"{prefix}gr", "table_perf": f"{prefix}perf",
        "table_monitoring": f"{prefix}monitor", "table_federation_global": f"{prefix}fedglobal",
        "table_federation_jobs": f"{prefix}fedjobs", "table_slurm": f"{prefix}slurm",
        "queue_federation": f"{prefix}fed", "queue_slurm": f"{prefix}slurm",
    }
    for k, v in names.items():
        if not (3 <= len(v) <= 63):
            raise ValueError(f"storage entity name '{v}' ({k}) must be 3-63 characters; shorten the prefix/pool id")
    return names


_SCHEMA = """
CREATE TABLE IF NOT EXISTS entities (kind TEXT NOT NULL, pk TEXT NOT NULL, rk TEXT NOT NULL, data TEXT NOT NULL,
    etag TEXT NOT NULL, updated REAL NOT NULL, PRIMARY KEY (kind, pk, rk));
CREATE TABLE IF NOT EXISTS queue (id INTEGER PRIMARY KEY AUTOINCREMENT, queue TEXT NOT NULL, body TEXT NOT NULL,
    visible_at REAL NOT NULL, dequeue_count INTEGER NOT NULL DEFAULT 0, pop_receipt TEXT, inserted REAL NOT NULL);
CREATE INDEX IF NOT EXISTS queue_idx ON queue(queue, visible_at);
CREATE TABLE IF NOT EXISTS leases (name TEXT PRIMARY KEY, holder TEXT NOT NULL, expires REAL NOT NULL);
CREATE TABLE IF NOT EXISTS events (id INTEGER PRIMARY KEY AUTOINCREMENT, ts REAL NOT NULL, pool TEXT, node TEXT,
    source TEXT NOT NULL, event TEXT NOT NULL, message TEXT);
"""


class Store:
    def __init__(self, root: Optional[str] = None):
        self.root = os.path.abspath(root or default_state_dir())
        os.makedirs(self.root, exist_ok=True)
        os.makedirs(os.path.join(self.root, "blobs"), exist_ok=True)
        self.db_path = os.path.join(self.root, "state.db")
        self._local = threading.local()
        c = self._conn()
        for stmt in _SCHEMA.split(";"):
            if stmt.strip():
                c.execute(stmt)

    # -- connection handling ---------------------------------------------------
    def _conn(self) -> sqlite3.Connection:
        c = getattr(self._local, "conn", None)
        if c is None:
            c = sqlite3.connect(self.db_path, timeout=60.0, isolation_level=None)
            c.execute("PRAGMA journal_mode=WAL")
            c.execute("PRAGMA synchronous=NORMAL")
            c.execute("PRAGMA busy_timeout=60000")
            self._local.conn = c
        return c

    @contextmanager
    def _tx(self):
        c = self._conn()
        c.execute("BEGIN IMMEDIATE")
        try:
            yield c
            c.execute("COMMIT")
        except BaseException:
            c.execute("ROLLBACK")
            raise

    def close(self) -> None:
        c = getattr(self._local, "conn", None)
        if c is not None:
            c.close()
            self._local.conn = None

    # -- table entities ----------------------------------------------------------
    @staticmethod
    def _row(r) -> dict:
        d = json.loads(r[3])
        d["_kind"], d["_pk"], d["_rk"], d["_etag"], d["_updated"] = r[0], r[1], r[2], r[4], r[5]
        return d

    @staticmethod
    def _clean(data: dict) -> str:
        return json.dumps({k: v for k, v in data.items() if not k.startswith("_")}, sort_keys=True, default=str)

    def insert(self, kind: str, pk: str, rk: str, data: dict, replace: bool = False) -> str:
        etag = uuid.uuid4().hex
        with self._tx() as c:
            if not replace and c.execute("SELECT 1 FROM entities WHERE kind=? AND pk=? AND rk=?", (kind, pk, rk)).fetchone():
                raise EntityExists(f"{kind}/{pk}/{rk}")
            c.execute("INSERT OR REPLACE INTO entities VALUES (?,?,?,?,?,?)", (kind, pk, rk, self._clean(data), etag, time.time()))
        return etag

    def insert_many(self, kind: str, pk: str, rows: list, replace: bool = False) -> int:
        """Insert ``[(rk, data), ...]`` in ONE transaction (a task collection): all rows or none; EntityExists names the first clash."""
        now = time.time()
        with self._tx() as c:
            if not replace:
                for rk, _ in rows:
                    if c.execute("SELECT 1 FROM entities WHERE kind=? AND pk=? AND rk=?", (kind, pk, rk)).fetchone():
                        raise EntityExists(f"{kind}/{pk}/{rk}")
            c.executemany("INSERT OR REPLACE INTO entities VALUES (?,?,?,?,?,?)",
                          [(kind, pk, rk, self._clean(data), uuid.uuid4().hex, now) for rk, data in rows])
        return len(rows)

    def get(self, kind: str, pk: str, rk: str) -> dict:
        r = self._conn().execute("SELECT kind,pk,rk,data,etag,updated FROM entities WHERE kind=? AND pk=? AND rk=?",
                                 (kind, pk, rk)).fetchone()
        if r is None:
            raise NotFound(f"{kind}/{pk}/{rk}")
        return self._row(r)

    def try_get(self, kind: str, pk: str, rk: str) -> Optional[dict]:
        try:
            return self.get(kind, pk, rk)
        except NotFound:
            return None

    def query(self, kind: str, pk: Optional[str] = None, rk_prefix: Optional[str] = None) -> list[dict]:
        q, args = "SELECT kind,pk,rk,data,etag,updated FROM entities WHERE kind=?", [kind]
        if pk is not None:
            q += " AND pk=?"; args.append(pk)
        if rk_prefix is not None:
            q += " AND rk LIKE ?"; args.append(rk_prefix.replace("%", r"\%") + "%")
        q += " ORDER BY pk, rk"
        return [self._row(r) for r in self._conn().execute(q, args).fetchall()]

    def update(self, kind: str, pk: str, rk: str, data: dict, etag: Optional[str] = None) -> str:
        """Replace; with ``etag`` this is a compare-and-swap (raises EtagMismatch)."""
        new = uuid.uuid4().hex
        with self._tx() as c:
            cur = c.execute("SELECT etag FROM entities WHERE kind=? AND pk=? AND rk=?", (kind, pk, rk)).fetchone()
            if cur is None:
                raise NotFound(f"{kind}/{pk}/{rk}")
            if etag is not None and cur[0] != etag:
                raise EtagMismatch(f"{kind}/{pk}/{rk}")
            c.execute("UPDATE entities SET data=?, etag=?, updated=? WHERE kind=? AND pk=? AND rk=?",
                      (self._clean(data), new, time.time(), kind, pk, rk))
        return new

    def merge(self, kind: str, pk: str, rk: str, patch: dict, create: bool = False) -> dict:
        """Read-modify-write under one transaction (shallow merge)."""
        with self._tx() as c:
            r = c.execute("SELECT kind,pk,rk,data,etag,updated FROM entities WHERE kind=? AND pk=? AND rk=?", (kind, pk, rk)).fetchone()
            if r is None:
                if not create:
                    raise NotFound(f"{kind}/{pk}/{rk}")
                cur: dict = {}
            else:
                cur = json.loads(r[3])
            cur.update({k: v for k, v in patch.items() if not k.startswith("_")})
            c.execute("INSERT OR REPLACE INTO entities VALUES (?,?,?,?,?,?)",
                      (kind, pk, rk, self._clean(cur), uuid.uuid4().hex, time.time()))
        return cur

    def mutate(self, kind: str, pk: str, rk: str, fn) -> dict:
        """Atomically apply ``fn(dict) -> dict|None`` to an entity."""
        with self._tx() as c:
            r = c.execute("SELECT data FROM entities WHERE kind=? AND pk=? AND rk=?", (kind, pk, rk)).fetchone()
            if r is None:
                raise NotFound(f"{kind}/{pk}/{rk}")
            cur = json.loads(r[0])
            out = fn(cur)
            cur = cur if out is None else out
            c.execute("UPDATE entities SET data=?, etag=?, updated=? WHERE kind=? AND pk=? AND rk=?",
                      (self._clean(cur), uuid.uuid4().hex, time.time(), kind, pk, rk))
        return cur

    def delete(self, kind: str, pk: Optional[str] = None, rk: Optional[str] = None) -> int:
        q, args = "DELETE FROM entities WHERE kind=?", [kind]
        if pk is not None:
            q += " AND pk=?"; args.append(pk)
        if rk is not None:
            q += " AND rk=?"; args.append(rk)
        with self._tx() as c:
            return c.execute(q, args).rowcount

    def exists(self, kind: str, pk: str, rk: str) -> bool:
        return self.try_get(kind, pk, rk) is not None

    # -- queues ------------------------------------------------------------------
    def put_message(self, queue: str, body: Any, delay: float = 0.0) -> int:
        with self._tx() as c:
            cur = c.execute("INSERT INTO queue(queue, body, visible_at, inserted) VALUES (?,?,?,?)",
                            (queue, json.dumps(body, default=str), time.time() + delay, time.time()))
            return int(cur.lastrowid)

    def get_messages(self, queue: str, n: int = 1, visibility_timeout: float = 30.0) -> list[dict]:
        now = time.time()
        out = []
        with self._tx() as c:
            rows = c.execute("SELECT id, body, dequeue_count FROM queue WHERE queue=? AND visible_at<=? ORDER BY id LIMIT ?",
                             (queue, now, n)).fetchall()
            for mid, body, dq in rows:
                receipt = uuid.uuid4().hex
                c.execute("UPDATE queue SET visible_at=?, dequeue_count=?, pop_receipt=? WHERE id=?",
                          (now + visibility_timeout, dq + 1, receipt, mid))
                out.append({"id": mid, "body": json.loads(body), "dequeue_count": dq + 1, "pop_receipt": receipt})
        return out

    def peek_messages(self, queue: str, n: int = 32) -> list[dict]:
        rows = self._conn().execute("SELECT id, body, dequeue_count FROM queue WHERE queue=? ORDER BY id LIMIT ?", (queue, n)).fetchall()
        return [{"id": r[0], "body": json.loads(r[1]), "dequeue_count": r[2]} for r in rows]

    def delete_message(self, queue: str, mid: int, pop_receipt: Optional[str] = None) -> bool:
        with self._tx() as c:
            if pop_receipt is None:
                return c.execute("DELETE FROM queue WHERE queue=? AND id=?", (queue, mid)).rowcount > 0
            return c.execute("DELETE FROM queue WHERE queue=? AND id=? AND pop_receipt=?", (queue, mid, pop_receipt)).rowcount > 0

    def clear_queue(self, queue: str) -> int:
        with self._tx() as c:
            return c.execute("DELETE FROM queue WHERE queue=?", (queue,)).rowcount

    def queue_length(self, queue: str) -> int:
        return int(self._conn().execute("SELECT COUNT(*) FROM queue WHERE queue=?", (queue,)).fetchone()[0])

    # -- leases (TTL locks) --------------------------------------------------------
    def acquire_lease(self, name: str, holder: str, duration: float) -> bool:
        now = time.time()
        with self._tx() as c:
            r = c.execute("SELECT holder, expires FROM leases WHERE name=?", (name,)).fetchone()
            if r is not None and r[1] > now and r[0] != holder:
                return False
            c.execute("INSERT OR REPLACE INTO leases VALUES (?,?,?)", (name, holder, now + duration))
            return True

    def renew_lease(self, name: str, holder: str, duration: float) -> bool:
        now = time.time()
        with self._tx() as c:
            return c.execute("UPDATE leases SET expires=? WHERE name=? AND holder=? AND expires>?",
                             (now + duration, name, holder, now)).rowcount > 0

    def release_lease(self, name: str, holder: str) -> bool:
        with self._tx() as c:
            return c.execute("DELETE FROM leases WHERE name=? AND holder=?", (name, holder)).rowcount > 0

    def lease_holder(self, name: str) -> Optional[str]:
        r = self._conn().execute("SELECT holder, expires FROM leases WHERE name=?", (name,)).fetchone()
        return r[0] if r and r[1] > time.time() else None

    # -- blobs -----------------------------------------------------------------------
    def blob_path(self, container: str, name: str) -> str:
        p = os.path.normpath(os.path.join(self.root, "blobs", container, name))
        if not p.startswith(os.path.join(self.root, "blobs") + os.sep):
            raise ValueError("blob name escapes the store")
        return p

    def put_blob(self, container: str, name: str, data: bytes) -> str:
        p = self.blob_path(container, name)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        tmp = p + f".tmp.{os.getpid()}.{uuid.uuid4().hex[:6]}"
        with open(tmp, "wb") as f:
            f.write(data); f.flush(); os.fsync(f.fileno())
        os.replace(tmp, p)      # atomic publish
        return p

    def put_blob_from_file(self, container: str, name: str, src: str) -> str:
        p = self.blob_path(container, name)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        tmp = p + f".tmp.{os.getpid()}"
        shutil.copyfile(src, tmp)
        os.replace(tmp, p)
        return p

    def get_blob(self, container: str, name: str) -> bytes:
        with open(self.blob_path(container, name), "rb") as f:
            return f.read()

    def list_blobs(self, container: str, prefix: str = "") -> list[str]:
        base = os.path.join(self.root, "blobs", container)
        out = []
        for d, _, fs in os.walk(base):
            for fn in fs:
                if ".tmp." in fn:
                    continue
                rel = os.path.relpath(os.path.join(d, fn), base).replace(os.sep, "/")
                if rel.startswith(prefix):
                    out.append(rel)
        return sorted(out)

    def delete_blob(self, container: str, name: str) -> bool:
        try:
            os.remove(self.blob_path(container, name)); return True
        except FileNotFoundError:
            return False

    def delete_container(self, container: str) -> None:
        shutil.rmtree(os.path.join(self.root, "blobs", container), ignore_errors=True)

    # -- timing / perf events ------------------------------------------------------------
    def record_event(self, source: str, event: str, pool: Optional[str] = None, node: Optional[str] = None,
                     message: Optional[str] = None, ts: Optional[float] = None) -> None:
        with self._tx() as c:
            c.execute("INSERT INTO events(ts, pool, node, source, event, message) VALUES (?,?,?,?,?,?)",
                      (ts if ts is not None else time.time(), pool, node, source, event, message))

    def events(self, pool: Optional[str] = None) -> list[dict]:
        q, args = "SELECT ts, pool, node, source, event, message FROM events", []
        if pool is not None:
            q += " WHERE pool=?"; args.append(pool)
        q += " ORDER BY ts, id"
        return [{"ts": r[0], "pool": r[1], "node": r[2], "source": r[3], "event": r[4], "message": r[5]}
                for r in self._conn().execute(q, args).fetchall()]

    def clear_events(self, pool: Optional[str] = None) -> None:
        with self._tx() as c:
            if pool is None:
                c.execute("DELETE FROM events")
            else:
                c.execute("DELETE FROM events WHERE pool=?", (pool,))

    # -- lifecycle ---------------------------------------------------------------------
    def clear_all(self) -> None:
        with self._tx() as c:
            for t in ("entities", "queue", "leases", "events"):
                c.execute(f"DELETE FROM {t}")
        shutil.rmtree(os.path.join(self.root, "blobs"), ignore_errors=True)
        os.makedirs(os.path.join(self.root, "blobs"), exist_ok=True)
