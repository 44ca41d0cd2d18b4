"""Storage clusters and managed disks on the box (the ``fs`` verbs).

The reference provisions managed disks and NFS/GlusterFS server VMs, RAID/btrfs-formats the
disks and exports them (/root/reference/convoy/remotefs.py:56-2040, scripts/
shipyard_remotefs_bootstrap.sh).  On one box a "disk" is a backing directory (on NVMe or
tmpfs), a "storage cluster" is a directory striped over its disks via symlinked bricks, and
pools mount it as ``storage_cluster`` shared data volumes.  Same verbs and lifecycle states
(add/orchestrate/resize/expand/suspend/start/status/del) so configs and scripts keep working.
"""
from __future__ import annotations

import os
import shutil
import time
from typing import Optional

from ..config import settings as S


class RemoteFsError(RuntimeError):
    pass


def _root(b) -> str:
    d = os.path.join(b.root, "remotefs")
    os.makedirs(d, exist_ok=True)
    return d


def _disk_dir(b, name: str) -> str:
    return os.path.join(_root(b), "disks", name)


def create_disks(b, config: dict) -> dict:
    md = S.remotefs_managed_disks(config)
    made = []
    for name in md["disk_names"]:
        d = _disk_dir(b, name)
        new = not os.path.isdir(d)
        os.makedirs(d, exist_ok=True)
        b.store.insert("disk", "local", name, {"name": name, "size_gb": md["disk_size_gb"], "sku": md["sku"], "path": d,
                                               "created": time.time(), "attached_to": None}, replace=new or True)
        made.append(name)
    return {"disks": made, "sku": md["sku"], "disk_size_gb": md["disk_size_gb"]}


def list_disks(b) -> list:
    return [{"name": d["name"], "size_gb": d["size_gb"], "sku": d["sku"], "attached_to": d.get("attached_to"), "path": d["path"]}
            for d in b.store.query("disk", "local")]


def delete_disks(b, config: dict, name: Optional[str] = None, all: bool = False) -> dict:
    names = [d["name"] for d in b.store.query("disk", "local")] if all else ([name] if name else S.remotefs_managed_disks(config)["disk_names"])
    deleted = []
    for n in names:
        d = b.store.try_get("disk", "local", n)
        if d is None:
            continue
        if d.get("attached_to"):
            raise RemoteFsError(f"disk {n} is attached to storage cluster {d['attached_to']}")
        shutil.rmtree(d["path"], ignore_errors=True)
        b.store.delete("disk", "local", n)
        deleted.append(n)
    return {"deleted": deleted}


def _cluster_path(b, sc: S.StorageCluster) -> str:
    return sc.local_path or os.path.join(_root(b), "clusters", sc.id)


def create_cluster(b, config: dict, cluster_id: str) -> dict:
    scs = S.remotefs_storage_clusters(config)
    if cluster_id not in scs:
        raise RemoteFsError(f"storage cluster {cluster_id} is not defined in the fs config")
    sc = scs[cluster_id]
    if b.store.exists("storagecluster", "local", cluster_id):
        raise RemoteFsError(f"storage cluster {cluster_id} already exists")
    path = _cluster_path(b, sc)
    os.makedirs(path, exist_ok=True)
    bricks = []
    for vm, m in (sc.raw.get("vm_disk_map") or {}).items():
        for dn in m.get("disk_array") or []:
            d = b.store.try_get("disk", "local", dn)
            if d is None:
                raise RemoteFsError(f"disk {dn} does not exist; run `fs disks add` (or `fs cluster orchestrate`)")
            if d.get("attached_to") not in (None, cluster_id):
                raise RemoteFsError(f"disk {dn} is already attached to {d['attached_to']}")
            b.store.merge("disk", "local", dn, {"attached_to": cluster_id})
            brick = os.path.join(path, ".bricks", f"vm{vm}-{dn}")
            os.makedirs(os.path.dirname(brick), exist_ok=True)
            if not os.path.lexists(brick):
                os.symlink(d["path"], brick)
            bricks.append({"vm": int(vm), "disk": dn, "filesystem": m.get("filesystem"), "raid_level": m.get("raid_level")})
    rec = {"id": cluster_id, "state": "running", "type": sc.file_server_type, "mountpoint": sc.mountpoint, "path": path,
           "vm_count": sc.vm_count, "bricks": bricks, "mount_options": sc.mount_options, "created": time.time(),
           "samba": (sc.raw.get("file_server") or {}).get("samba")}
    b.store.insert("storagecluster", "local", cluster_id, rec)
    return rec


def _get(b, cluster_id: str) -> dict:
    rec = b.store.try_get("storagecluster", "local", cluster_id)
    if rec is None:
        raise RemoteFsError(f"storage cluster {cluster_id} does not exist")
    return rec


def resize_cluster(b, config: dict, cluster_id: str) -> dict:
    sc = S.remotefs_storage_clusters(config).get(cluster_id)
    rec = _get(b, cluster_id)
    if sc is None:
        raise RemoteFsError(f"storage cluster {cluster_id} is not defined in the fs config")
    if rec["type"] != "glusterfs":
        raise RemoteFsError("only glusterfs storage clusters can be resized")
    if sc.vm_count < rec["vm_count"]:
        raise RemoteFsError("storage clusters can only grow")
    b.store.merge("storagecluster", "local", cluster_id, {"vm_count": sc.vm_count})
    return {"id": cluster_id, "vm_count": sc.vm_count}


def expand_cluster(b, config: dict, cluster_id: str, rebalance: bool = True) -> dict:
    sc = S.remotefs_storage_clusters(config).get(cluster_id)
    rec = _get(b, cluster_id)
    have = {x["disk"] for x in rec["bricks"]}
    added = []
    for vm, m in ((sc.raw.get("vm_disk_map") if sc else None) or {}).items():
        for dn in m.get("disk_array") or []:
            if dn in have:
                continue
            d = b.store.try_get("disk", "local", dn)
            if d is None:
                raise RemoteFsError(f"disk {dn} does not exist")
            b.store.merge("disk", "local", dn, {"attached_to": cluster_id})
            brick = os.path.join(rec["path"], ".bricks", f"vm{vm}-{dn}")
            os.makedirs(os.path.dirname(brick), exist_ok=True)
            if not os.path.lexists(brick):
                os.symlink(d["path"], brick)
            rec["bricks"].append({"vm": int(vm), "disk": dn, "filesystem": m.get("filesystem"), "raid_level": m.get("raid_level")})
            added.append(dn)
    b.store.merge("storagecluster", "local", cluster_id, {"bricks": rec["bricks"]})
    return {"id": cluster_id, "added_disks": added, "rebalanced": bool(rebalance and added)}


def set_cluster_state(b, cluster_id: str, state: str) -> dict:
    _get(b, cluster_id)
    b.store.merge("storagecluster", "local", cluster_id, {"state": state})
    return {"id": cluster_id, "state": state}


def cluster_status(b, cluster_id: str, detail: bool = False) -> dict:
    rec = _get(b, cluster_id)
    out = {"id": cluster_id, "state": rec["state"], "type": rec["type"], "path": rec["path"], "mountpoint": rec["mountpoint"],
           "vm_count": rec["vm_count"], "disks": [x["disk"] for x in rec["bricks"]]}
    if detail and os.path.isdir(rec["path"]):
        st = shutil.disk_usage(rec["path"])
        nfiles = sum(len(fs) for _, _, fs in os.walk(rec["path"]))
        out.update({"bytes_total": st.total, "bytes_free": st.free, "files": nfiles, "bricks": rec["bricks"]})
    return out


def delete_cluster(b, cluster_id: str, delete_data: bool = False) -> dict:
    rec = _get(b, cluster_id)
    for x in rec["bricks"]:
        if b.store.exists("disk", "local", x["disk"]):
            b.store.merge("disk", "local", x["disk"], {"attached_to": None})
    if delete_data:
        shutil.rmtree(rec["path"], ignore_errors=True)
    b.store.delete("storagecluster", "local", cluster_id)
    return {"deleted": True, "id": cluster_id, "data_deleted": delete_data}


def mount_args_for_pool(b, cluster_id: str) -> dict:
    """What a pool needs to bind the cluster as a ``storage_cluster`` shared data volume
    (the reference builds an fstab line, remotefs.py:56)."""
    rec = _get(b, cluster_id)
    fs = "nfs4" if rec["type"] == "nfs" else "glusterfs"
    return {"fstab": f"127.0.0.1:{rec['mountpoint']} {rec['path']} {fs} {','.join(rec.get('mount_options') or ['defaults'])} 0 0",
            "host_path": rec["path"]}
