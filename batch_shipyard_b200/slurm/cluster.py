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
from typing import Optional

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


# ---- daemon / slurmctld entry points ---------------------------------------------------------------------------------
def expand_hostlist(spec: str) -> list[str]:
    """Slurm hostlist expression -> names: ``c-p-pool-[0-2,5],login0`` (what slurmctld passes to Resume/SuspendProgram)."""
    out, depth, cur = [], 0, ""
    for ch in spec:
        if ch == "[":
            depth += 1
        elif ch == "]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur:
        out.append(cur)
    names = []
    for item in out:
        if "[" not in item:
            names.append(item); continue
        prefix, rest = item.split("[", 1)
        body, suffix = rest.rsplit("]", 1)
        for part in body.split(","):
            if "-" in part:
                lo, hi = part.split("-", 1)
                width = len(lo)
                names += [f"{prefix}{str(i).zfill(width)}{suffix}" for i in range(int(lo), int(hi) + 1)]
            else:
                names.append(f"{prefix}{part}{suffix}")
    return names


def process_retry_queue(b, config: dict, max_messages: int = 32) -> dict:
    """One daemon pass: re-drive resumes that failed earlier (bounded by MAX_RESUME_FAILURE_ATTEMPTS)."""
    cid = S.slurm_options(config)["cluster_id"]
    q = f"slurm-retry-{cid}"
    retried, dropped = [], []
    for msg in b.store.get_messages(q, n=max_messages, visibility_timeout=30.0):
        body = msg["body"]
        host = b.store.try_get("slurmhost", cid, body["host"])
        b.store.delete_message(q, msg["id"], msg.get("pop_receipt"))
        if host is None or host.get("state") in ("up", "provisioning_error"):
            dropped.append(body["host"]); continue
        r = resume(b, config, [body["host"]])
        (retried if r["resumed"] else dropped).append(body["host"])
    return {"retried": retried, "dropped": dropped}


def daemon(b, config: dict, poll_interval: float = 5.0, max_iterations: int = 0) -> dict:
    """Long-running companion of slurmctld (reference: `slurm.py daemon`): drains the retry queue every poll interval."""
    it, total = 0, {"retried": 0, "dropped": 0}
    while max_iterations <= 0 or it < max_iterations:
        r = process_retry_queue(b, config)
        total["retried"] += len(r["retried"]); total["dropped"] += len(r["dropped"])
        it += 1
        if max_iterations <= 0 or it < max_iterations:
            time.sleep(poll_interval)
    return dict(total, iterations=it)


def node_assignment(b, config: dict, node_or_host: str) -> Optional[dict]:
    """The Slurm host bound to a GPU node id (or the record of a host name): what a compute node asks for when it comes up
    (/root/reference/slurm/slurm.py `get-node-assignment`)."""
    cid = S.slurm_options(config)["cluster_id"]
    for h in b.store.query("slurmhost", cid):
        if h.get("gpu_node") == node_or_host or h.get("name") == node_or_host or h.get("_rk") == node_or_host:
            return {"host": h.get("name") or h.get("_rk"), "node": h.get("gpu_node"), "state": h.get("state"),
                    "partition": h.get("partition"), "pool": h.get("pool"), "assignment_complete": bool(h.get("assignment_complete"))}
    return None


def main(argv=None) -> int:
    """``python -m batch_shipyard_b200.slurm.cluster <verb> --conf slurm.yaml [--hosts LIST | --hostfile FILE | --host NAME]``
    — the programs a slurm.conf generated by ``slurm_conf`` points ResumeProgram / SuspendProgram / ResumeFailProgram at, plus the
    node-side verbs of the reference's helper (/root/reference/slurm/slurm.py:1448-1466: daemon, sakey, resume, resume-fail, suspend,
    check-provisioning-status, get-node-assignment, complete-node-assignment; ``--hostfile`` lines are ``host [partition]``)."""
    import argparse
    import json
    from ..backend.local import LocalBackend
    from ..config.loader import load_file
    ap = argparse.ArgumentParser(prog="shipyard-slurm")
    ap.add_argument("verb", choices=["resume", "suspend", "resume-fail", "daemon", "sakey", "check-provisioning-status",
                                     "get-node-assignment", "complete-node-assignment"])
    ap.add_argument("--conf", required=True)
    ap.add_argument("--hosts", default="")
    ap.add_argument("--hostfile")
    ap.add_argument("--host")
    ap.add_argument("--state-dir", default=os.environ.get("SHIPYARD_STATE_DIR"))
    ap.add_argument("--poll-interval", type=float, default=5.0)
    ap.add_argument("--iterations", type=int, default=0)
    a = ap.parse_args(argv)
    config = load_file(a.conf)
    b = LocalBackend(state_dir=a.state_dir) if a.state_dir else LocalBackend()
    hosts = expand_hostlist(a.hosts) if a.hosts else []
    if a.hostfile:
        with open(a.hostfile) as f:
            for line in f:
                if line.split():
                    hosts += expand_hostlist(line.split()[0])
    if a.host and a.verb in ("resume", "suspend", "resume-fail"):
        hosts += expand_hostlist(a.host)
    if a.verb == "resume":
        out = resume(b, config, hosts)
    elif a.verb == "suspend":
        out = suspend(b, config, hosts)
    elif a.verb == "resume-fail":
        out = resume_failed(b, config, hosts)
    elif a.verb == "daemon":
        out = daemon(b, config, a.poll_interval, a.iterations)
    elif a.verb == "sakey":
        out = {"storage_account": "local", "key": None, "note": "the state store is local: no storage account key is needed"}
    else:
        if not a.host:
            ap.error(f"{a.verb} needs --host")
        rec = node_assignment(b, config, a.host)
        if rec is None:
            print(json.dumps({"host": a.host, "assigned": False}))
            return 1
        if a.verb == "complete-node-assignment":
            b.store.merge("slurmhost", S.slurm_options(config)["cluster_id"], rec["host"], {"assignment_complete": True})
            rec["assignment_complete"] = True
        out = dict(rec, assigned=True)
        if a.verb == "check-provisioning-status" and rec["state"] != "up":
            print(json.dumps(out))
            return 1
    print(json.dumps(out))
    return 0 if not out.get("failed") else 1


if __name__ == "__main__":
    import sys
    sys.exit(main())
