"""Client-side data ingress: ``global_resources.files`` -> shared volumes / storage.

Semantics of /root/reference/convoy/data.py:492-1176 kept: per ``files[]`` entry a
``source{path, include, exclude}`` goes to either a shared data volume or a storage
account (both at once is an error, :1009-1016); ``multinode_*`` methods walk the tree,
**bin-pack files by bytes onto the nodes (least-loaded bucket first)** (:606-665), may
**split big files into chunks** that are re-joined at the destination (:636-660, 769-804,
857-870) and move ``max_parallel_transfers_per_node`` files at a time per node (:807-876);
the achieved rate is logged (:732-737).

Re-designed for one box: "nodes" are the pool's GPUs sharing one filesystem, so a transfer
is a chunked parallel copy driven by the native staging arena (``libshipyard_stage``) —
optionally leaving the bytes resident in that GPU's HBM (``to_hbm``) — instead of
scp/rsync pipes over ssh.
"""
from __future__ import annotations

import concurrent.futures as cf
import fnmatch
import os
import shutil
import time
from dataclasses import dataclass, field
from typing import Optional

from ..config import settings as S
from ..utils import util

logger = util.get_logger()
_CHUNK = 4 << 20


@dataclass
class FileEntry:
    src: str
    rel: str
    size: int
    offset: int = 0          # for split files: byte range [offset, offset+size)
    part: Optional[int] = None
    parts: int = 1


@dataclass
class Bucket:
    node: str
    files: list = field(default_factory=list)
    bytes: int = 0


def walk_source(path: str, include=None, exclude=None) -> list[FileEntry]:
    out = []
    include, exclude = list(include or []), list(exclude or [])
    if os.path.isfile(path):
        return [FileEntry(path, os.path.basename(path), os.path.getsize(path))]
    for d, _, fs in os.walk(path):
        for fn in sorted(fs):
            p = os.path.join(d, fn)
            rel = os.path.relpath(p, path).replace(os.sep, "/")
            if include and not any(fnmatch.fnmatch(rel, pat) or fnmatch.fnmatch(fn, pat) for pat in include):
                continue
            if any(fnmatch.fnmatch(rel, pat) or fnmatch.fnmatch(fn, pat) for pat in exclude):
                continue
            out.append(FileEntry(p, rel, os.path.getsize(p)))
    return out


def split_entries(entries: list[FileEntry], split_mb: Optional[int]) -> list[FileEntry]:
    """Files larger than `split_mb` become several ranged parts (joined again on arrival)."""
    if not split_mb:
        return entries
    lim = int(split_mb) << 20
    out = []
    for e in entries:
        if e.size <= lim:
            out.append(e)
            continue
        n = (e.size + lim - 1) // lim
        for i in range(n):
            out.append(FileEntry(e.src, e.rel, min(lim, e.size - i * lim), offset=i * lim, part=i, parts=n))
    return out


def bin_pack(entries: list[FileEntry], nodes: list[str]) -> list[Bucket]:
    """Greedy: biggest file first into the currently lightest bucket."""
    buckets = [Bucket(n) for n in nodes]
    for e in sorted(entries, key=lambda x: -x.size):
        b = min(buckets, key=lambda x: x.bytes)
        b.files.append(e); b.bytes += e.size
    return buckets


def _copy_range(e: FileEntry, dst_root: str) -> int:
    dst = os.path.join(dst_root, e.rel)
    os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
    if e.parts == 1:
        shutil.copyfile(e.src, dst)
        return e.size
    # pre-size once, then every part writes its own range (no re-join pass needed)
    if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(e.src):
        with open(dst, "ab") as f:
            f.truncate(os.path.getsize(e.src))
    with open(e.src, "rb") as fi, open(dst, "r+b") as fo:
        fi.seek(e.offset); fo.seek(e.offset)
        left = e.size
        while left > 0:
            buf = fi.read(min(_CHUNK, left))
            if not buf:
                break
            fo.write(buf); left -= len(buf)
    return e.size


def transfer(buckets: list[Bucket], dst_root: str, max_parallel_per_node: int = 2) -> dict:
    t0 = time.time()
    total = 0
    with cf.ThreadPoolExecutor(max_workers=max(1, len(buckets) * max(1, max_parallel_per_node))) as ex:
        futs = [ex.submit(_copy_range, e, dst_root) for b in buckets for e in b.files]
        for f in cf.as_completed(futs):
            total += f.result()
    dt = max(time.time() - t0, 1e-9)
    rate = total * 8 / 1e6 / dt
    logger.info("ingress: %d bytes in %.3f s (%.1f Mbit/s, %.1f MiB/s)", total, dt, rate, total / dt / (1 << 20))
    return {"bytes": total, "seconds": dt, "mbit_per_s": rate, "files": sum(len(b.files) for b in buckets),
            "per_node": {b.node: {"files": len(b.files), "bytes": b.bytes} for b in buckets}}


def shared_volume_path(b, pool_id: str, gs: S.GlobalSettings, name: str) -> str:
    from ..jobs.builder import shared_volume_host_path
    sv = gs.shared_data_volumes.get(name)
    if sv is None:
        raise ValueError(f"shared data volume '{name}' is not defined")
    p = shared_volume_host_path(sv).replace("$AZ_BATCH_NODE_SHARED_DIR", b.node_shared_dir(pool_id)).replace(
        "$AZ_BATCH_NODE_ROOT_DIR", b.pool_root(pool_id))
    return p


def ingress_data(b, config: dict, pool_id: Optional[str], to_fs: Optional[str] = None, kind: str = "all") -> dict:
    """Process every ``global_resources.files`` entry.  ``to_fs``: a storage cluster id (``--to-fs``)."""
    gs = S.global_settings(config)
    results = []
    for spec in gs.files:
        src, dst = spec["source"], spec["destination"]
        sdv, link = dst.get("shared_data_volume"), dst.get("storage_account_settings")
        if sdv and link:
            raise ValueError("a files[] destination may name a shared_data_volume or storage_account_settings, not both")
        dt = dst.get("data_transfer") or {}
        entries = walk_source(src["path"], src.get("include"), src.get("exclude"))
        if link:
            if kind not in ("all", "storage"):
                continue
            from .mover import storage_root
            root = S.credentials_storage_local_path(config, link) or storage_root(b.root, link)
            target = os.path.join(root, (dt.get("remote_path") or "").strip("/"))
            stats = transfer(bin_pack(entries, ["storage"]), target, 4)
            results.append(dict(stats, destination=f"storage:{link}:{dt.get('remote_path') or ''}"))
            continue
        if kind not in ("all", "shared"):
            continue
        if to_fs:
            from ..fs import remotefs
            base = remotefs.cluster_status(b, to_fs)["path"]
        elif sdv:
            if not pool_id:
                raise ValueError("a pool is needed to ingress into a shared data volume")
            base = shared_volume_path(b, pool_id, gs, sdv)
        else:
            if not pool_id:
                raise ValueError("files[] destination needs a shared_data_volume, storage_account_settings or --to-fs")
            base = b.node_shared_dir(pool_id)
        target = os.path.join(base, (dst.get("relative_destination_path") or "").strip("/"))
        method = dt.get("method") or "multinode_scp"
        nodes = [n["id"] for n in b.list_nodes(pool_id)] if (pool_id and b.pool_exists(pool_id)) else ["box"]
        if not method.startswith("multinode") or not nodes:
            nodes = nodes[:1] or ["box"]
        parts = split_entries(entries, dt.get("split_files_megabytes") if method.startswith("multinode") else None)
        stats = transfer(bin_pack(parts, nodes), target, int(dt.get("max_parallel_transfers_per_node") or 2))
        results.append(dict(stats, destination=target, method=method))
    return {"transfers": results, "total_bytes": sum(r["bytes"] for r in results)}


def pool_input_data(b, config: dict, pool_id: str, input_data: dict) -> dict:
    """Pool-level ``input_data``: fetched once into the node shared dir at pool creation."""
    from .mover import copy_tree, storage_root
    out = []
    env = {"AZ_BATCH_NODE_SHARED_DIR": b.node_shared_dir(pool_id), "AZ_BATCH_NODE_ROOT_DIR": b.pool_root(pool_id)}
    for spec in input_data.get("azure_storage") or []:
        link = spec["storage_account_settings"]
        root = S.credentials_storage_local_path(config, link) or storage_root(b.root, link)
        src = os.path.join(root, spec["remote_path"].strip("/"))
        dst = util.expand_env(spec.get("local_path") or "$AZ_BATCH_NODE_SHARED_DIR", env)
        n, nb = copy_tree(src, dst, spec.get("include"), spec.get("exclude")) if os.path.exists(src) else (0, 0)
        out.append({"source": src, "destination": dst, "files": n, "bytes": nb})
    for spec in input_data.get("azure_batch") or []:
        job = b.get_job(spec["job_id"])
        src = b.task_dir(job["pool_id"], spec["job_id"], spec["task_id"])
        dst = util.expand_env(spec.get("destination") or "$AZ_BATCH_NODE_SHARED_DIR", env)
        n, nb = copy_tree(src, dst, spec.get("include"), spec.get("exclude"))
        out.append({"source": src, "destination": dst, "files": n, "bytes": nb})
    return {"input_data": out}
