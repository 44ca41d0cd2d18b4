"""Task-side data mover (the blobxfer / task_file_mover stand-in).

Runs inside a task's system prologue / epilogue:
  fetch      copy one resource file (file:// URL or path) into the working dir
  ingress    storage "account" (a local directory) -> task directory, include/exclude filters
  egress     task directory -> storage location, only when the condition matches
             $SHIPYARD_TASK_RESULT (tasksuccess | taskfailure | taskcompletion)
  taskfiles  copy another task's files (``input_data.azure_batch``)
Reference: /root/reference/scripts/shipyard_blobxfer.sh:7-71 (spec parsing, conditional
egress), /root/reference/cargo/task_file_mover.py:87-141, /root/reference/convoy/data.py:115-489.
"""
from __future__ import annotations

import argparse
from typing import Optional
import fnmatch
import os
import shutil
import sys
import time
import urllib.parse
from typing import Iterable


def _match(rel: str, include: Iterable[str], exclude: Iterable[str]) -> bool:
    include, exclude = list(include or []), list(exclude or [])
    if include and not any(fnmatch.fnmatch(rel, p) for p in include):
        return False
    return not any(fnmatch.fnmatch(rel, p) for p in exclude)


_NATIVE_MIN_BYTES = 1 << 20          # smaller files: the per-ticket overhead is not worth it, shutil.copy2 in this process


class _Copier:
    """Bounded-concurrency file copier: files of 1 MB and more go to the worker threads of the native staging library (copy_file_range
    in the kernel, mode + mtime preserved: ``sy_stage_submit_copy``) and run in parallel; small files are copied inline.  The reference moves task
    data with blobxfer's parallel transfers (/root/reference/convoy/data.py:219-291); ``SHIPYARD_NATIVE_COPY=0`` forces the inline path."""

    def __init__(self):
        self.stager, self.tickets = None, []
        if os.environ.get("SHIPYARD_NATIVE_COPY", "1") not in ("0", "", "off", "false"):
            try:
                from ..ops.stage import Stager
                self.stager = Stager(None, arena_bytes=64 << 20, concurrency=int(os.environ.get("SHIPYARD_COPY_THREADS", "4") or 4))
            except Exception as e:  # noqa: BLE001 - the library is optional for the mover: say so once, then copy inline
                _log("download", f"native copier unavailable ({type(e).__name__}: {e}); copying inline")

    def copy(self, src: str, out: str) -> None:
        if self.stager is not None and os.path.getsize(src) >= _NATIVE_MIN_BYTES:
            self.tickets.append((self.stager.submit_copy(src, out), src))
        else:
            shutil.copy2(src, out)

    def finish(self) -> None:
        """Wait for every queued copy; the first failure is raised after all tickets have been collected."""
        err = None
        for t, src in self.tickets:
            try:
                self.stager.wait(t)
            except Exception as e:  # noqa: BLE001
                err = err or OSError(f"copy of {src} failed: {e}")
            finally:
                self.stager.release(t)
        self.tickets = []
        if self.stager is not None:
            self.stager.close()
            self.stager = None
        if err:
            raise err


def copy_tree(src: str, dst: str, include=(), exclude=(), collect: Optional[list] = None) -> tuple[int, int]:
    """Copy files under `src` (or the single file `src`) to `dst`; returns (files, bytes); `collect` receives the written paths."""
    n = nb = 0
    cp = _Copier()
    try:
        if os.path.isfile(src):
            os.makedirs(dst, exist_ok=True)
            out = os.path.join(dst, os.path.basename(src))
            cp.copy(src, out)
            if collect is not None:
                collect.append(out)
            return 1, os.path.getsize(src)
        for d, _, fs in os.walk(src):
            for fn in fs:
                p = os.path.join(d, fn)
                rel = os.path.relpath(p, src).replace(os.sep, "/")
                if not _match(rel, include, exclude):
                    continue
                out = os.path.join(dst, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                cp.copy(p, out)
                if collect is not None:
                    collect.append(out)
                n += 1; nb += os.path.getsize(p)
        return n, nb
    finally:
        cp.finish()


def record_stage_manifest(paths: list, source: str) -> Optional[str]:
    """Append ingressed files to the task's staging manifest ($SHIPYARD_STAGE_MANIFEST, set by the runner spec): the list the
    task-side stager (ops.stage.stage_task_inputs) pushes file -> pinned arena -> HBM while the first step runs."""
    mpath = os.environ.get("SHIPYARD_STAGE_MANIFEST")
    if not mpath or not paths:
        return None
    import json
    try:
        with open(mpath) as f:
            man = json.load(f)
    except (OSError, ValueError):
        man = {"version": 1, "files": []}
    known = {e["path"] for e in man["files"]}
    for p in paths:
        if p not in known:
            man["files"].append({"path": p, "bytes": os.path.getsize(p), "source": source})
    tmp = mpath + ".tmp"
    with open(tmp, "w") as f:
        json.dump(man, f, indent=1)
    os.replace(tmp, mpath)
    return mpath


def storage_root(state_dir: str, link: str) -> str:
    """Directory backing a storage account link: credentials.storage.<link>.local_path if recorded
    at pool creation, else <state>/storage/<link>."""
    from ..state.store import Store
    st = Store(state_dir)
    ent = st.try_get("storagelink", link, "")
    if ent and ent.get("local_path"):
        return ent["local_path"]
    return os.path.join(state_dir, "storage", link)


def remote_path(root: str, remote: str) -> str:
    """`container/prefix` inside a storage root; a path that climbs out of the root (`..`) is refused — a storage account is a
    directory here, and a job file must not be able to read or overwrite the rest of the box through it."""
    p = os.path.normpath(os.path.join(root, remote.strip("/")))
    base = os.path.normpath(root)
    if p != base and not p.startswith(base + os.sep):
        raise ValueError(f"remote path '{remote}' leaves the storage account directory")
    return p


def _log(name: str, msg: str) -> None:
    """Per-task transfer log (the reference's blobxfer-{download,upload}.log, scripts/shipyard_blobxfer.sh:69).  Written only inside a
    task directory: outside a task (tests, ad-hoc CLI use) there is no log file, so nothing lands in the current directory."""
    tdir = os.environ.get("AZ_BATCH_TASK_DIR")
    if not tdir or not os.path.isdir(tdir):
        return
    try:
        with open(os.path.join(tdir, f"blobxfer-{name}.log"), "a") as f:
            f.write(f"{time.strftime('%Y-%m-%dT%H:%M:%S')} {msg}\n")
    except OSError:
        pass


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="shipyard-mover")
    sub = ap.add_subparsers(dest="cmd", required=True)
    f = sub.add_parser("fetch"); f.add_argument("--url", required=True); f.add_argument("--dest", required=True); f.add_argument("--mode")
    for name in ("ingress", "egress"):
        p = sub.add_parser(name)
        p.add_argument("--state-dir", required=True); p.add_argument("--link", required=True)
        p.add_argument("--remote", required=True); p.add_argument("--local", required=True)
        p.add_argument("--include", action="append", default=[]); p.add_argument("--exclude", action="append", default=[])
        if name == "egress":
            p.add_argument("--condition", default="tasksuccess", choices=["tasksuccess", "taskfailure", "taskcompletion"])
    t = sub.add_parser("taskfiles")
    t.add_argument("--state-dir", required=True); t.add_argument("--job", required=True); t.add_argument("--task", required=True)
    t.add_argument("--dest", required=True)
    t.add_argument("--include", action="append", default=[]); t.add_argument("--exclude", action="append", default=[])
    a = ap.parse_args(argv)
    if a.cmd == "fetch":
        u = urllib.parse.urlparse(a.url)
        if u.scheme in ("", "file"):
            src = u.path if u.scheme == "file" else a.url
        else:
            print(f"mover: cannot fetch '{a.url}': no network on a local pool (use file:// or a path)", file=sys.stderr)
            return 1
        os.makedirs(os.path.dirname(os.path.abspath(a.dest)) or ".", exist_ok=True)
        shutil.copy2(src, a.dest)
        if a.mode:
            os.chmod(a.dest, int(str(a.mode), 8))
        return 0
    if a.cmd == "ingress":
        try:
            src = remote_path(storage_root(a.state_dir, a.link), a.remote)
        except ValueError as e:
            print(f"mover: {e}", file=sys.stderr)
            return 1
        if not os.path.exists(src):
            print(f"mover: ingress source {src} does not exist", file=sys.stderr)
            return 1
        written: list = []
        n, nb = copy_tree(src, a.local, a.include, a.exclude, collect=written)
        record_stage_manifest(written, f"{a.link}:{a.remote}")
        _log("download", f"ingress {a.link}:{a.remote} -> {a.local}: {n} files, {nb} bytes")
        return 0
    if a.cmd == "egress":
        result = os.environ.get("SHIPYARD_TASK_RESULT", "success")
        want = {"tasksuccess": result == "success", "taskfailure": result != "success", "taskcompletion": True}[a.condition]
        if not want:
            _log("upload", f"egress skipped: condition {a.condition} not met (result={result})")
            return 0
        try:
            dst = remote_path(storage_root(a.state_dir, a.link), a.remote)
        except ValueError as e:
            print(f"mover: {e}", file=sys.stderr)
            return 1
        n, nb = copy_tree(a.local, dst, a.include, a.exclude + ["*.spec", ".shipyard.envlist", ".heartbeat"])
        _log("upload", f"egress {a.local} -> {a.link}:{a.remote}: {n} files, {nb} bytes")
        return 0
    if a.cmd == "taskfiles":
        from ..backend.local import LocalBackend
        b = LocalBackend(state_dir=a.state_dir)
        job = b.get_job(a.job)
        src = b.task_dir(job["pool_id"], a.job, a.task)
        n, nb = copy_tree(src, a.dest, a.include, a.exclude)
        _log("download", f"taskfiles {a.job}/{a.task} -> {a.dest}: {n} files, {nb} bytes")
        return 0
    return 2


if __name__ == "__main__":
    sys.exit(main())
