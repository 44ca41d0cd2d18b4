"""Federation client side: federations, member pools, job submission and action queues.

Reference: /root/reference/convoy/storage.py:679-1745 (federation create/destroy, pool
add/remove, job sequence packing with etag CAS, queue messages), /root/reference/convoy/
batch.py:5900-6109 (v1 message + constraints metadata), /root/reference/convoy/federation.py
(proxy provisioning -> here: start/stop the local scheduler daemon).

Wire format v1 kept: queue message ``{version:'1', federation_id, target, blob_data, uuid}``;
blob = serialised ``{version, action{method, kind}, <kind>{id, data, constraints, task_naming},
task_map}`` (JSON here instead of pickle: no arbitrary-code deserialisation).
"""
from __future__ import annotations

import dataclasses
import json
import os
import signal
import subprocess
import sys
import time
import uuid
from typing import Optional

from ..config import settings as S
from ..state.store import EtagMismatch
from ..utils import util
from . import constraints as FC


class FederationError(RuntimeError):
    pass


def _fid(federation_id: str) -> str:
    return federation_id.lower()


def _queue(federation_id: str) -> str:
    return "fed-" + util.hash_federation_id(federation_id)


def create_federation(b, federation_id: str, force: bool = False, unique_job_ids: bool = True) -> dict:
    fid = _fid(federation_id)
    if b.store.exists("federation", fid, "") and not force:
        raise FederationError(f"federation {fid} already exists (use --force)")
    b.store.insert("federation", fid, "", {"id": fid, "hash": util.hash_federation_id(fid), "unique_job_ids": unique_job_ids,
                                           "created": time.time()}, replace=True)
    return {"federation_id": fid, "hash": util.hash_federation_id(fid), "unique_job_ids": unique_job_ids}


def _require(b, federation_id: str) -> dict:
    f = b.store.try_get("federation", _fid(federation_id), "")
    if f is None:
        raise FederationError(f"federation {federation_id} does not exist")
    return f


def list_federations(b, ids: Optional[list] = None) -> dict:
    out = {}
    for f in b.store.query("federation"):
        if ids and f["id"] not in [_fid(i) for i in ids]:
            continue
        out[f["id"]] = {"hash": f["hash"], "unique_job_ids": f["unique_job_ids"],
                        "pools": [p["_rk"] for p in b.store.query("fedpool", f["id"])]}
    return out


def destroy_federation(b, federation_id: str) -> dict:
    fid = _fid(federation_id)
    _require(b, fid)
    b.store.delete("fedpool", fid); b.store.delete("fedjob", fid); b.store.delete("fedblocked", fid); b.store.delete("fedseq", fid)
    b.store.clear_queue(_queue(fid))
    b.store.delete_container("fed-" + util.hash_federation_id(fid))
    b.store.delete("federation", fid, "")
    return {"destroyed": True, "federation_id": fid}


def add_pools(b, federation_id: str, pools: list) -> dict:
    fid = _fid(federation_id)
    _require(b, fid)
    for p in pools:
        if not b.pool_exists(p):
            raise FederationError(f"pool {p} does not exist")
        b.store.insert("fedpool", fid, p, {"added": time.time(), "batch_service_url": "local"}, replace=True)
    return list_federations(b, [fid])


def remove_pools(b, federation_id: str, pools: list, all_: bool = False) -> dict:
    fid = _fid(federation_id)
    _require(b, fid)
    if all_:
        b.store.delete("fedpool", fid)
    for p in pools:
        b.store.delete("fedpool", fid, p)
    return list_federations(b, [fid])


def _append_sequence(b, fid: str, target: str, unique_id: str) -> int:
    """FIFO per target job: append the action's unique id with compare-and-swap on the sequence entity."""
    for _ in range(50):
        ent = b.store.try_get("fedseq", fid, target)
        if ent is None:
            try:
                b.store.insert("fedseq", fid, target, {"sequence": [unique_id]})
                return 1
            except Exception:  # noqa: BLE001 - lost the creation race, retry as an update
                continue
        seq = list(ent["sequence"]) + [unique_id]
        try:
            b.store.update("fedseq", fid, target, {"sequence": seq}, etag=ent["_etag"])
            return len(seq)
        except EtagMismatch:
            continue
    raise FederationError("could not append to the federation job sequence (contention)")


def submit_job_to_federation(b, config: dict, federation_id: str, job_rec: dict, task_records: list, jobspec: dict,
                             recurrence=None) -> dict:
    fid = _fid(federation_id)
    fed = _require(b, fid)
    kind = "job_schedule" if recurrence is not None else "job"
    cons = FC.parse_constraints(jobspec, task_records)
    if recurrence is not None:
        cons.task.tasks_per_recurrence = len(task_records)   # (the reference never sets this: SURVEY.md Q12)
    if fed["unique_job_ids"] and b.store.exists("fedjob", fid, job_rec["id"]):
        raise FederationError(f"job {job_rec['id']} already exists in federation {fid} (unique job ids are required)")
    uid = str(uuid.uuid4())
    gid = S.global_settings(config).autogenerated_task_id
    jauto = S.autogenerated_task_id(jobspec.get("autogenerated_task_id"), gid)
    payload = {"version": "1", "action": {"method": "add", "kind": kind},
               kind: {"id": job_rec["id"], "data": job_rec, "constraints": dataclasses.asdict(cons),
                      "task_naming": {"prefix": jauto.prefix, "padding": jauto.zfill_width},
                      "recurrence_interval_s": recurrence.interval.total_seconds() if recurrence else None},
               "task_map": task_records}
    container = "fed-" + fed["hash"]
    blob = f"messages/{uid}.json"
    b.store.put_blob(container, blob, json.dumps(payload, default=str).encode())
    _append_sequence(b, fid, job_rec["id"], uid)
    b.store.put_message(_queue(fid), {"version": "1", "federation_id": fid, "target": job_rec["id"], "blob_data": f"{container}/{blob}", "uuid": uid})
    out = {"federation": {"id": fid, "storage": container}, "kind": kind, "action": "add", "unique_id": uid}
    out["tasks_per_recurrence" if recurrence else "num_tasks"] = len(task_records)
    return out


def enqueue_job_action(b, federation_id: str, method: str, targets: list, all_: bool = False) -> dict:
    fid = _fid(federation_id)
    fed = _require(b, fid)
    if all_:
        targets = [j["_rk"] for j in b.store.query("fedjob", fid)]
    out = {}
    for t in targets:
        uid = str(uuid.uuid4())
        container = "fed-" + fed["hash"]
        blob = f"messages/{uid}.json"
        b.store.put_blob(container, blob, json.dumps({"version": "1", "action": {"method": method, "kind": "job"}, "job": {"id": t}}).encode())
        _append_sequence(b, fid, t, uid)
        b.store.put_message(_queue(fid), {"version": "1", "federation_id": fid, "target": t, "blob_data": f"{container}/{blob}", "uuid": uid})
        out[t] = {"action": method, "unique_id": uid}
    return out


def list_jobs(b, federation_id: str, blocked: bool = False, queued: bool = False, job_id: Optional[str] = None) -> dict:
    fid = _fid(federation_id)
    _require(b, fid)
    if blocked:
        return {"blocked": [{"unique_id": x["_rk"], "target": x.get("target"), "reason": x.get("reason"), "since": x.get("since")}
                            for x in b.store.query("fedblocked", fid) if not job_id or x.get("target") == job_id]}
    if queued:
        return {"queued": [{"unique_id": m["body"]["uuid"], "target": m["body"]["target"], "dequeue_count": m["dequeue_count"]}
                           for m in b.store.peek_messages(_queue(fid), 1000) if not job_id or m["body"]["target"] == job_id]}
    return {"jobs": {j["_rk"]: {"pool_id": j.get("pool_id"), "kind": j.get("kind"), "unique_id": j.get("unique_id"), "scheduled": j.get("scheduled")}
                     for j in b.store.query("fedjob", fid) if not job_id or j["_rk"] == job_id}}


def zap_action(b, federation_id: str, unique_id: str) -> dict:
    fid = _fid(federation_id)
    _require(b, fid)
    removed = 0
    for m in b.store.peek_messages(_queue(fid), 10000):
        if m["body"].get("uuid") == unique_id:
            b.store.delete_message(_queue(fid), m["id"]); removed += 1
    removed += b.store.delete("fedblocked", fid, unique_id)
    for s in b.store.query("fedseq", fid):
        if unique_id in s["sequence"]:
            b.store.update("fedseq", fid, s["_rk"], {"sequence": [u for u in s["sequence"] if u != unique_id]})
    return {"zapped": unique_id, "removed": removed}


# -- daemon lifecycle ("proxy" verbs) ----------------------------------------------------------
def proxy_create(b, config: dict, state_dir: str) -> dict:
    st = b.store.try_get("service", "fedproxy", "") or {}
    try:
        if st.get("pid"):
            os.kill(int(st["pid"]), 0)
            return dict(proxy_status(b), note="already running")
    except OSError:
        pass
    po = S.federation_proxy_options(config)
    d = os.path.join(state_dir, "federation")
    os.makedirs(d, exist_ok=True)
    log = open(os.path.join(d, po.log_filename), "ab")
    p = subprocess.Popen([sys.executable, "-m", "batch_shipyard_b200.fed.daemon", "--state-dir", state_dir,
                          "--federations-interval", str(po.federations_polling_interval), "--actions-interval", str(po.actions_polling_interval),
                          "--blackout", str(po.scheduling_after_success_blackout_interval), "--log-level", po.log_level]
                         + (["--evaluate-autoscale"] if po.scheduling_after_success_evaluate_autoscale else []),
                         stdout=log, stderr=log, stdin=subprocess.DEVNULL, start_new_session=True,
                         cwd=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    log.close()                               # the daemon holds its own copy of the descriptor
    b.store.insert("service", "fedproxy", "", {"pid": p.pid, "state": "running", "started": time.time(), "log": os.path.join(d, po.log_filename)}, replace=True)
    return proxy_status(b)


def proxy_status(b) -> dict:
    st = b.store.try_get("service", "fedproxy", "")
    if not st:
        return {"state": "absent"}
    alive = False
    try:
        if st.get("pid"):
            os.kill(int(st["pid"]), 0); alive = True
    except OSError:
        pass
    return {"state": "running" if alive else st.get("state", "dead") if st.get("state") == "suspended" else ("running" if alive else "dead"),
            "pid": st.get("pid"), "leader": b.store.lease_holder("federation-leader"), "log": st.get("log")}


def proxy_stop(b, destroy: bool = False) -> dict:
    st = b.store.try_get("service", "fedproxy", "")
    if st and st.get("pid"):
        try:
            os.kill(int(st["pid"]), signal.SIGTERM)
        except OSError:
            pass
    if destroy:
        b.store.delete("service", "fedproxy", "")
        return {"destroyed": True}
    if st:
        b.store.merge("service", "fedproxy", "", {"state": "suspended", "pid": None})
    return proxy_status(b)
