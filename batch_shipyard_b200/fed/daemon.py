"""Federation scheduler daemon ("proxy").

Reference behaviour (/root/reference/federation/federation.py): leader election through a
15 s lease renewed every 5 s so several daemons can run for HA (:962-986), poll the
federation table and each federation's action queue (:3135-3185), enforce FIFO per job
through the sequence entity (:1242-1330, 3029-3066), evaluate constraints and pick a pool by
greedy best fit (:2030-2228), create the job + tasks on the chosen pool (:2410-2656), keep
unschedulable actions in a *blocked* table instead of dropping them (:1332-1364, 2341-2360),
black out a pool for a while after scheduling onto it and force an autoscale evaluation
(:577-607, 948-957).
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import sys
import time
import uuid

from ..backend.agent import spawn_detached_agent
from ..backend.local import BackendError, LocalBackend
from . import constraints as FC
from .scheduler import select_pool

LEADER_LEASE_S, LEADER_RENEW_S = 15.0, 5.0
log = logging.getLogger("fedproxy")


def _cons_from_dict(d: dict) -> FC.Constraints:
    return FC.Constraints(FC.PoolConstraints(**d["pool"]), FC.ComputeNodeConstraints(**d["compute_node"]), FC.TaskConstraints(**d["task"]))


class FederationProcessor:
    def __init__(self, b: LocalBackend, blackout: float = 15.0, evaluate_autoscale: bool = True, holder: str = ""):
        self.b, self.blackout, self.evaluate_autoscale = b, blackout, evaluate_autoscale
        self.holder = holder or f"fedproxy-{os.getpid()}-{uuid.uuid4().hex[:6]}"
        self.blackouts: dict = {}
        self._last_renew = 0.0

    # -- leader election -------------------------------------------------------------------
    def is_leader(self) -> bool:
        now = time.time()
        if now - self._last_renew < LEADER_RENEW_S and self.b.store.lease_holder("federation-leader") == self.holder:
            return True
        ok = self.b.store.acquire_lease("federation-leader", self.holder, LEADER_LEASE_S)
        if ok:
            self._last_renew = now
        return ok

    # -- one pass over every federation queue --------------------------------------------------
    def process_all(self) -> int:
        n = 0
        for fed in self.b.store.query("federation"):
            n += self.process_federation(fed)
        return n

    def process_federation(self, fed: dict) -> int:
        fid, q = fed["id"], "fed-" + fed["hash"]
        handled = 0
        for msg in self.b.store.get_messages(q, n=32, visibility_timeout=30.0):
            body = msg["body"]
            target, uid = body["target"], body["uuid"]
            seq = self.b.store.try_get("fedseq", fid, target)
            if seq is None or uid not in seq["sequence"]:
                self.b.store.delete_message(q, msg["id"], msg["pop_receipt"])      # zapped or stale
                continue
            if seq["sequence"][0] != uid:
                continue        # not at the head of this job's FIFO yet; becomes visible again later
            try:
                container, _, name = body["blob_data"].partition("/")
                payload = json.loads(self.b.store.get_blob(container, name))
            except Exception as e:  # noqa: BLE001
                log.error("dropping unreadable action %s: %s", uid, e)
                self._finish(fid, q, msg, target, uid)
                continue
            done, reason = self.apply_action(fid, target, uid, payload)
            if done:
                self.b.store.delete("fedblocked", fid, uid)
                self._finish(fid, q, msg, target, uid)
                handled += 1
            else:
                self.b.store.insert("fedblocked", fid, uid, {"target": target, "reason": reason, "since": time.time()}, replace=True)
        return handled

    def _finish(self, fid: str, q: str, msg: dict, target: str, uid: str) -> None:
        self.b.store.delete_message(q, msg["id"], msg["pop_receipt"])
        ent = self.b.store.try_get("fedseq", fid, target)
        if ent is not None:
            rest = [u for u in ent["sequence"] if u != uid]
            if rest:
                self.b.store.update("fedseq", fid, target, {"sequence": rest})
            else:
                self.b.store.delete("fedseq", fid, target)

    # -- actions ---------------------------------------------------------------------------------
    def apply_action(self, fid: str, target: str, uid: str, payload: dict) -> tuple[bool, str]:
        method, kind = payload["action"]["method"], payload["action"]["kind"]
        if method == "add":
            return self.add_job(fid, uid, payload, kind)
        fj = self.b.store.try_get("fedjob", fid, target)
        if fj is None:
            return True, "unknown job"      # nothing to do: consume the action
        try:
            if fj.get("kind") == "job_schedule":
                (self.b.delete_job_schedule if method == "delete" else self.b.terminate_job_schedule)(target)
            elif self.b.job_exists(target):
                (self.b.delete_job if method == "delete" else self.b.terminate_job)(target)
        except BackendError as e:
            return False, str(e)
        if method == "delete":
            self.b.store.delete("fedjob", fid, target)
        return True, ""

    def add_job(self, fid: str, uid: str, payload: dict, kind: str) -> tuple[bool, str]:
        spec = payload[kind]
        cons = _cons_from_dict(spec["constraints"])
        now = time.time()
        views = []
        for fp in self.b.store.query("fedpool", fid):
            try:
                views.append(FC.pool_view(self.b, fp["_rk"], self.blackouts.get(fp["_rk"], 0.0)))
            except BackendError:
                continue
        if not views:
            return False, "federation has no pools"
        pool_id, diag = select_pool(views, cons, now)
        if pool_id is None:
            return False, "; ".join(f"{k}: {v}" for k, v in sorted(diag.items())) or "no pool satisfies the constraints"
        job = dict(spec["data"], pool_id=pool_id)
        tasks = [dict(t) for t in payload.get("task_map") or []]
        # a job prepared for one pool shape may land on another: drop stale placement
        for t in tasks:
            t.pop("node_ids", None)
        try:
            if kind == "job_schedule":
                self.b.add_job_schedule({"id": job["id"], "pool_id": pool_id, "recurrence_interval_s": spec.get("recurrence_interval_s") or 60.0,
                                         "job_template": {k: v for k, v in job.items() if k != "id"}, "task_map": tasks})
            else:
                if not self.b.job_exists(job["id"]):
                    self.b.add_job(job)
                existing = self.b.task_ids(job["id"])
                naming = spec.get("task_naming") or {}
                from ..jobs.builder import next_generic_task_id
                renamed = {}
                for t in tasks:
                    if t["id"] in existing:       # regenerate colliding generic ids, keep dependencies consistent
                        new = next_generic_task_id(existing, naming.get("prefix", "task-"), int(naming.get("padding", 5)))
                        renamed[t["id"]] = new; t["id"] = new
                    existing.add(t["id"])
                for t in tasks:
                    t["depends_on"] = [renamed.get(d, d) for d in t.get("depends_on") or []]
                self.b.add_tasks(job["id"], tasks)
        except BackendError as e:
            return False, str(e)
        self.b.store.insert("fedjob", fid, job["id"], {"pool_id": pool_id, "kind": kind, "unique_id": uid, "scheduled": now,
                                                        "why": diag.get(pool_id)}, replace=True)
        self.blackouts[pool_id] = now + self.blackout
        if self.evaluate_autoscale:
            pool = self.b.get_pool(pool_id)
            if (pool.get("autoscale") or {}).get("enabled") and (pool.get("autoscale") or {}).get("formula"):
                try:
                    from ..backend.agent import NodeAgent
                    NodeAgent(self.b, pool_id).evaluate_autoscale(apply=True)
                except Exception as e:  # noqa: BLE001
                    log.warning("autoscale evaluation on %s failed: %s", pool_id, e)
        if not os.environ.get("SHIPYARD_FED_NO_AGENT"):
            spawn_detached_agent(self.b.root, pool_id)
        log.info("scheduled %s %s of federation %s on pool %s (%s)", kind, job["id"], fid, pool_id, diag.get(pool_id))
        return True, ""


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="shipyard federation scheduler daemon")
    ap.add_argument("--state-dir", required=True)
    ap.add_argument("--federations-interval", type=float, default=15.0)
    ap.add_argument("--actions-interval", type=float, default=5.0)
    ap.add_argument("--blackout", type=float, default=15.0)
    ap.add_argument("--evaluate-autoscale", action="store_true")
    ap.add_argument("--log-level", default="info")
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args(argv)
    logging.basicConfig(level=getattr(logging, a.log_level.upper(), logging.INFO), format="%(asctime)s %(levelname)s %(name)s - %(message)s")
    b = LocalBackend(state_dir=a.state_dir)
    proc = FederationProcessor(b, a.blackout, a.evaluate_autoscale)
    while True:
        if proc.is_leader():
            proc.process_all()
        if a.once:
            return 0
        time.sleep(max(0.2, a.actions_interval))


if __name__ == "__main__":
    sys.exit(main())
