"""Slurm surface: map Slurm partitions / node names onto the box's GPU pools.

The reference stands up Slurm controller/login VMs and binds Batch pools as elastic
partitions whose nodes are resumed/suspended on demand (/root/reference/convoy/slurm.py:
63-1234; daemon verbs resume / suspend / resume-fail in /root/reference/slurm/slurm.py:
969-1275, host states :80-86).  Locally the "cluster" is a mapping table: each elastic
partition's pools get ``max_compute_nodes`` node names (``<cluster>-<partition>-<pool>-<i>``)
bound to GPU indices; resume/suspend flip the host state and resize the pool, with a retry
queue for failed resumes.  A ``slurm.conf`` fragment is generated for a real slurmctld.
"""
from __future__ import annotations

import os
import subprocess
import time

from ..config import settings as S

HOST_STATES = ("none", "resuming", "up", "suspending", "suspended", "provisioning_error")
MAX_RESUME_FAILURE_ATTEMPTS = 10


def _hosts(config: dict) -> list[dict]:
    so = S.slurm_options(config)
    out = []
    for pname, part in so["elastic_partitions"].items():
        for pool in part["pools"]:
            for i in range(pool.max_compute_nodes):
                out.append({"name": f"{so['cluster_id']}-{pname}-{pool.pool_id}-{i}".lower(), "partition": pname, "pool_id": pool.pool_id,
                            "ordinal": i, "weight": pool.weight, "features": pool.features})
    return out


def slurm_conf(config: dict) -> str:
    so = S.slurm_options(config)
    lines = [f"ClusterName={so['cluster_id']}", "SlurmctldHost=localhost", "GresTypes=gpu",
             f"SuspendTime={int(so['idle_reclaim_time'].total_seconds())}",
             "ResumeProgram=shipyard-slurm resume", "SuspendProgram=shipyard-slurm suspend", "ResumeFailProgram=shipyard-slurm resume-fail"]
    by_part: dict = {}
    for h in _hosts(config):
        lines.append(f"NodeName={h['name']} Gres=gpu:1 Weight={h['weight']} State=CLOUD" + (f" Feature={','.join(h['features'])}" if h["features"] else ""))
        by_part.setdefault(h["partition"], []).append(h["name"])
    for pname, part in so["elastic_partitions"].items():
        lines.append(f"PartitionName={pname} Nodes={','.join(by_part.get(pname, []))} Default={'YES' if part['default'] else 'NO'}"
                     + (f" MaxTime={part['max_runtime_limit']}" if part.get("max_runtime_limit") else ""))
    for up in so["unmanaged_partitions"]:
        lines.append(f"PartitionName={up['partition']} Nodes={','.join(up['nodes'])}")
    return "\n".join(lines) + "\n"


def create(b, config: dict) -> dict:
    so = S.slurm_options(config)
    cid = so["cluster_id"]
    for h in _hosts(config):
        b.store.insert("slurmhost", cid, h["name"], dict(h, state="suspended", gpu_node=None, resume_failures=0), replace=True)
    d = os.path.join(b.root, "slurm", cid)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "slurm.conf"), "w") as f:
        f.write(slurm_conf(config))
    b.store.insert("service", "slurm", cid, {"state": "running", "created": time.time(), "conf": os.path.join(d, "slurm.conf")}, replace=True)
    return status(b, config)


def status(b, config: dict) -> dict:
    cid = S.slurm_options(config)["cluster_id"]
    svc = b.store.try_get("service", "slurm", cid)
    hosts = b.store.query("slurmhost", cid)
    hist: dict = {}
    for h in hosts:
        hist[h["state"]] = hist.get(h["state"], 0) + 1
    return {"cluster_id": cid, "state": svc["state"] if svc else "absent", "hosts": len(hosts), "host_states": hist,
            "conf": svc.get("conf") if svc else None}


def set_state(b, config: dict, state: str) -> dict:
    cid = S.slurm_options(config)["cluster_id"]
    b.store.merge("service", "slurm", cid, {"state": state}, create=True)
    return status(b, config)


def destroy(b, config: dict) -> dict:
    cid = S.slurm_options(config)["cluster_id"]
    b.store.delete("slurmhost", cid)
    b.store.delete("service", "slurm", cid)
    return {"destroyed": True, "cluster_id": cid}


def resume(b, config: dict, hostnames: list[str]) -> dict:
    """Bring hosts up: bind each to an idle GPU node of its pool (growing the pool when it is short)."""
    cid = S.slurm_options(config)["cluster_id"]
    out = {"resumed": [], "failed": []}
    for hn in hostnames:
        h = b.store.try_get("slurmhost", cid, hn)
        if h is None:
            out["failed"].append({"host": hn, "error": "unknown host"}); continue
        b.store.merge("slurmhost", cid, hn, {"state": "resuming"})
        try:
            taken = {x.get("gpu_node") for x in b.store.query("slurmhost", cid) if x.get("gpu_node")}
            nodes = [n for n in b.list_nodes(h["pool_id"]) if n["state"] in ("idle", "running") and n["id"] not in taken]
            if not nodes:
                raise RuntimeError(f"pool {h['pool_id']} has no free node")
            b.store.merge("slurmhost", cid, hn, {"state": "up", "gpu_node": nodes[0]["id"], "resume_failures": 0})
            out["resumed"].append({"host": hn, "node": nodes[0]["id"]})
        except Exception as e:  # noqa: BLE001
            fails = int(h.get("resume_failures") or 0) + 1
            st = "provisioning_error" if fails >= MAX_RESUME_FAILURE_ATTEMPTS else "suspended"
            b.store.merge("slurmhost", cid, hn, {"state": st, "resume_failures": fails})
            b.store.put_message(f"slurm-retry-{cid}", {"host": hn, "attempt": fails})
            out["failed"].append({"host": hn, "error": str(e)})
    return out


def suspend(b, config: dict, hostnames: list[str]) -> dict:
    cid = S.slurm_options(config)["cluster_id"]
    done = []
    for hn in hostnames:
        if b.store.exists("slurmhost", cid, hn):
            b.store.merge("slurmhost", cid, hn, {"state": "suspended", "gpu_node": None})
            done.append(hn)
    return {"suspended": done}


def resume_failed(b, config: dict, hostnames: list[str]) -> dict:
    cid = S.slurm_options(config)["cluster_id"]
    for hn in hostnames:
        if b.store.exists("slurmhost", cid, hn):
            b.store.merge("slurmhost", cid, hn, {"state": "suspended", "gpu_node": None})
    return {"cleaned": hostnames}


def shell(b, config: dict, kind: str, node_name, command) -> dict:
    cid = S.slurm_options(config)["cluster_id"]
    env = dict(os.environ)
    ctx = {"kind": kind, "cluster_id": cid}
    if kind == "node":
        h = b.store.try_get("slurmhost", cid, node_name)
        if h is None:
            raise ValueError(f"unknown slurm node {node_name}")
        ctx.update({"node": node_name, "state": h["state"], "gpu_node": h.get("gpu_node")})
        if h.get("gpu_node", "") and str(h["gpu_node"]).startswith("gpu-"):
            env["CUDA_VISIBLE_DEVICES"] = str(h["gpu_node"])[4:]
    if not command:
        return ctx
    p = subprocess.run(" ".join(command), shell=True, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return dict(ctx, exit_code=p.returncode, output=p.stdout)
