"""Local "Batch service": pools of GPUs, jobs, tasks, schedules — state only.

This is the service half the reference gets from Azure Batch (pool / node /
job / task objects and their state machines, /root/reference/convoy/batch.py:
85-100 node states, 625-830 allocation + recovery, 3884-3911 task counts).
A *node* is one GPU of the box (``gpu-<i>``) or, on a CPU-only box, a virtual
slot (``cpu-<i>``).  All methods are plain state transitions on the shared
store, so any number of CLI invocations, node agents and daemons can use it
concurrently.  Execution lives in ``backend.agent``.
"""
from __future__ import annotations

import datetime
import os
import re
import shutil
import signal
import time
from typing import Iterable, Optional

from .. import __version__
from ..config import settings as S
from ..state.store import EntityExists, NotFound, Store

NODE_STATES = ("creating", "idle", "leaving_pool", "offline", "preempted", "rebooting", "reimaging", "running",
               "start_task_failed", "starting", "unknown", "unusable", "waiting_for_start_task")
READY_STATES = ("idle", "preempted", "running")
TASK_STATES = ("active", "preparing", "running", "completed")
JOB_STATES = ("active", "disabling", "disabled", "enabling", "terminating", "completed", "deleting")
MAX_TASKS_PER_REQUEST = 100
MAX_REBOOT_RETRIES = 5


class BackendError(RuntimeError):
    pass


# Azure Batch enforces this server-side (the reference never had to); here ids become directory names and shell words, so the
# backend is the gate: /root/reference/convoy/batch.py hands ids straight to the service, which rejects anything else.
_ID_RE = re.compile(r"^[A-Za-z0-9_-]{1,64}$")


def validate_id(kind: str, value) -> str:
    if not isinstance(value, str) or not _ID_RE.match(value):
        raise BackendError(f"invalid {kind} id {value!r}: must match [A-Za-z0-9_-]{{1,64}}")
    return value


def proc_start_ticks(pid: int):
    """Kernel start time of `pid` (field 22 of /proc/<pid>/stat) or None: guards pid liveness checks against pid reuse."""
    try:
        with open(f"/proc/{int(pid)}/stat") as f:
            st = f.read()
        return int(st[st.rindex(")") + 2:].split()[19])
    except (OSError, ValueError, IndexError):
        return None


def pid_alive(pid, start_ticks=None) -> bool:
    """True while the process exists, is not a zombie and (when known) is still the process we started."""
    try:
        with open(f"/proc/{int(pid)}/stat") as f:
            st = f.read()
        fields = st[st.rindex(")") + 2:].split()
        if fields[0] in ("Z", "X"):
            return False
        return start_ticks is None or int(fields[19]) == int(start_ticks)
    except (OSError, ValueError, IndexError, TypeError):
        return False


def _now() -> float:
    return time.time()


def _iso(ts: Optional[float]) -> Optional[str]:
    return None if ts is None else datetime.datetime.fromtimestamp(ts, datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ")


def detect_gpus() -> list[dict]:
    """Enumerate GPUs through the native probe; empty on a CPU-only box."""
    from ..pool.topology import probe_gpus
    return probe_gpus()


class LocalBackend:
    def __init__(self, store: Optional[Store] = None, state_dir: Optional[str] = None):
        self.store = store or Store(state_dir)
        self.root = self.store.root

    # ------------------------------------------------------------------ dirs
    def pool_root(self, pool_id: str) -> str:
        return os.path.join(self.root, "pools", pool_id)

    def node_shared_dir(self, pool_id: str) -> str:
        return os.path.join(self.pool_root(pool_id), "shared")

    def node_startup_dir(self, pool_id: str) -> str:
        return os.path.join(self.pool_root(pool_id), "startup")

    def job_dir(self, pool_id: str, job_id: str) -> str:
        return os.path.join(self.pool_root(pool_id), "workitems", job_id, "job-1")

    def task_dir(self, pool_id: str, job_id: str, task_id: str) -> str:
        return os.path.join(self.job_dir(pool_id, job_id), task_id)

    # ------------------------------------------------------------------ pools
    def pool_exists(self, pool_id: str) -> bool:
        return self.store.exists("pool", pool_id, "")

    def get_pool(self, pool_id: str) -> dict:
        try:
            return self.store.get("pool", pool_id, "")
        except NotFound:
            raise BackendError(f"pool {pool_id} does not exist") from None

    def list_pools(self) -> list[dict]:
        return self.store.query("pool")

    def create_pool(self, ps: S.PoolSettings, gpus: Optional[list] = None, total_nodes: Optional[int] = None,
                    metadata: Optional[dict] = None) -> dict:
        """Create the pool object and its nodes in state ``creating`` (provisioning is the provisioner's job)."""
        validate_id("pool", ps.id)
        if self.pool_exists(ps.id):
            raise BackendError(f"pool {ps.id} already exists")
        gpus = list(gpus) if gpus is not None else []
        want = ps.vm_dedicated + ps.vm_low_priority
        if total_nodes is None:
            total_nodes = want if want > 0 else (len(gpus) if gpus else 1)
        if gpus and total_nodes > len(gpus):
            raise BackendError(f"pool {ps.id} wants {total_nodes} node(s) but the box has {len(gpus)} GPU(s): "
                               "vm_count (dedicated + low_priority) may not exceed the GPU count")
        now = _now()
        pool = {
            "id": ps.id, "state": "active", "allocation_state": "resizing", "vm_size": ps.vm_size,
            "target_dedicated": ps.vm_dedicated if want > 0 else total_nodes, "target_low_priority": ps.vm_low_priority,
            "max_tasks_per_node": ps.max_tasks_per_node, "node_fill_type": ps.node_fill_type,
            "inter_node_communication_enabled": ps.inter_node_communication_enabled, "native": ps.native,
            "created": now, "allocation_state_transition_time": now, "resize_errors": [],
            "autoscale": {"enabled": ps.autoscale is not None, "formula": None, "last_evaluation": None,
                          "evaluation_interval_s": ps.autoscale.evaluation_interval.total_seconds() if ps.autoscale else None},
            "gpus": gpus, "metadata": dict(metadata or {}, **{"BATCH_SHIPYARD_VERSION": __version__}),
            "per_job_auto_scratch": ps.per_job_auto_scratch, "reboot_on_start_task_failed": ps.reboot_on_start_task_failed,
            "attempt_recovery_on_unusable": ps.attempt_recovery_on_unusable,
        }
        self.store.insert("pool", ps.id, "", pool)
        for d in (self.pool_root(ps.id), self.node_shared_dir(ps.id), self.node_startup_dir(ps.id),
                  os.path.join(self.node_startup_dir(ps.id), "wd")):
            os.makedirs(d, exist_ok=True)
        n_ded = pool["target_dedicated"]
        for i in range(total_nodes):
            self._add_node(ps.id, i, gpus[i] if gpus else None, dedicated=i < n_ded)
        return pool

    def _add_node(self, pool_id: str, ordinal: int, gpu: Optional[int], dedicated: bool) -> dict:
        nid = f"gpu-{gpu}" if gpu is not None else f"cpu-{ordinal}"
        node = {"id": nid, "pool_id": pool_id, "ordinal": ordinal, "gpu_index": gpu, "state": "creating",
                "dedicated": dedicated, "state_transition_time": _now(), "allocation_time": _now(),
                "last_boot_time": None, "running_tasks": [], "total_tasks_run": 0, "total_tasks_succeeded": 0,
                "start_task": None, "errors": [], "reboots": 0, "scheduling": "enabled"}
        self.store.insert("node", pool_id, nid, node, replace=True)
        return node

    def set_node_state(self, pool_id: str, node_id: str, state: str, **extra) -> dict:
        if state not in NODE_STATES:
            raise BackendError(f"bad node state {state}")

        def fn(n):
            n["state"] = state
            n["state_transition_time"] = _now()
            n.update(extra)
        return self.store.mutate("node", pool_id, node_id, fn)

    def set_pool_allocation_state(self, pool_id: str, state: str) -> None:
        self.store.merge("pool", pool_id, "", {"allocation_state": state, "allocation_state_transition_time": _now()})

    def list_nodes(self, pool_id: str, start_task_failed: bool = False, unusable: bool = False) -> list[dict]:
        nodes = self.store.query("node", pool_id)
        if start_task_failed:
            nodes = [n for n in nodes if n["state"] == "start_task_failed"]
        if unusable:
            nodes = [n for n in nodes if n["state"] == "unusable"]
        return sorted(nodes, key=lambda n: n["ordinal"])

    def get_node(self, pool_id: str, node_id: str) -> dict:
        try:
            return self.store.get("node", pool_id, node_id)
        except NotFound:
            raise BackendError(f"node {node_id} does not exist in pool {pool_id}") from None

    def node_counts(self, pool_id: str) -> dict:
        self.get_pool(pool_id)
        out = {"dedicated": {s: 0 for s in NODE_STATES}, "low_priority": {s: 0 for s in NODE_STATES}}
        for n in self.list_nodes(pool_id):
            out["dedicated" if n["dedicated"] else "low_priority"][n["state"]] += 1
        for k in ("dedicated", "low_priority"):
            out[k]["total"] = sum(out[k][s] for s in NODE_STATES)
        return out

    def current_node_counts(self, pool_id: str) -> dict:
        nodes = self.list_nodes(pool_id)
        return {"current_dedicated": sum(1 for n in nodes if n["dedicated"]),
                "current_low_priority": sum(1 for n in nodes if not n["dedicated"])}

    def delete_pool(self, pool_id: str) -> None:
        validate_id("pool", pool_id)
        self.get_pool(pool_id)
        for j in self.list_jobs(pool_id=pool_id):
            if j["state"] not in ("completed", "deleting"):
                self.terminate_job(j["id"], reason="pool deleted")
        self.store.delete("node", pool_id)
        self.store.delete("pool", pool_id, "")
        self.store.delete("globalresource", pool_id)
        shutil.rmtree(self.pool_root(pool_id), ignore_errors=True)

    def resize_pool(self, pool_id: str, dedicated: int, low_priority: int) -> dict:
        """Grow (new nodes in ``creating``) or shrink (idle nodes first; busy ones go ``leaving_pool``)."""
        pool = self.get_pool(pool_id)
        nodes = self.list_nodes(pool_id)
        gpus = pool.get("gpus") or []
        total = dedicated + low_priority
        if gpus and total > len(gpus):
            raise BackendError(f"cannot resize pool {pool_id} to {total} nodes: box has {len(gpus)} GPUs")
        self.store.merge("pool", pool_id, "", {"target_dedicated": dedicated, "target_low_priority": low_priority,
                                                "allocation_state": "resizing", "allocation_state_transition_time": _now()})
        have = len(nodes)
        if total > have:
            used_gpus = {n["gpu_index"] for n in nodes}
            used_ord = {n["ordinal"] for n in nodes}
            free_gpus = [g for g in gpus if g not in used_gpus]
            nd = sum(1 for n in nodes if n["dedicated"])
            o = 0
            for _ in range(total - have):
                while o in used_ord:
                    o += 1
                used_ord.add(o)
                self._add_node(pool_id, o, free_gpus.pop(0) if gpus else None, dedicated=nd < dedicated)
                nd += 1
        elif total < have:
            # remove low-priority before dedicated, idle before busy, highest ordinal first
            order = sorted(nodes, key=lambda n: (n["dedicated"], n["state"] == "running", -n["ordinal"]))
            for n in order[: have - total]:
                if n["state"] == "running":
                    self.set_node_state(pool_id, n["id"], "leaving_pool")
                else:
                    self.store.delete("node", pool_id, n["id"])
        return self.get_pool(pool_id)

    def remove_node(self, pool_id: str, node_id: str) -> None:
        self.get_node(pool_id, node_id)
        self.store.delete("node", pool_id, node_id)
        c = self.current_node_counts(pool_id)
        self.store.merge("pool", pool_id, "", {"target_dedicated": c["current_dedicated"],
                                                "target_low_priority": c["current_low_priority"]})

    def pool_stats(self, pool_id: str) -> dict:
        pool = self.get_pool(pool_id)
        nodes = self.list_nodes(pool_id)
        now = _now()
        by_state: dict = {}
        for n in nodes:
            by_state[n["state"]] = by_state.get(n["state"], 0) + 1
        ready = [n for n in nodes if n.get("last_boot_time")]
        a2r = [n["last_boot_time"] - n["allocation_time"] for n in ready if n.get("allocation_time")]
        slots = len(nodes) * pool["max_tasks_per_node"]
        running = sum(len(n["running_tasks"]) for n in nodes)

        def agg(vals, with_sum=False):
            vals = [v for v in vals if v is not None]
            out = {"min": min(vals, default=None), "mean": (sum(vals) / len(vals)) if vals else None, "max": max(vals, default=None)}
            if with_sum:
                out["sum"] = sum(vals)
            return out

        # the fields `pool stats` of the reference prints (/root/reference/convoy/batch.py:1460-1633): node-state histogram, uptime,
        # allocation-to-ready and last start-task time, total / running tasks per node, scheduling slots busy / available / runnable
        runnable_nodes = [n for n in nodes if n["state"] in ("idle", "running")]
        runnable = len(runnable_nodes) * pool["max_tasks_per_node"]
        busy = sum(len(n["running_tasks"]) for n in runnable_nodes)
        return {"pool_id": pool_id, "node_states": by_state, "total_nodes": len(nodes),
                "dedicated_nodes": sum(1 for n in nodes if n["dedicated"]),
                "low_priority_nodes": sum(1 for n in nodes if not n["dedicated"]),
                "allocation_state": pool["allocation_state"], "created": _iso(pool["created"]),
                "uptime_s": agg(now - n["last_boot_time"] for n in ready),
                "allocation_to_ready_s": agg(a2r),
                "start_task_s": agg(n.get("start_task_seconds") for n in nodes),
                "total_tasks_run": agg((n.get("total_tasks_run") or 0 for n in nodes), with_sum=True),
                "running_tasks_per_node": agg((len(n["running_tasks"]) for n in nodes), with_sum=True),
                "running_tasks": running, "task_slots": slots,
                "scheduling_slots": {"busy": busy, "available": max(runnable - busy, 0), "runnable": runnable, "total": slots},
                "slot_utilization_pct": round(100.0 * running / slots, 2) if slots else 0.0}

    # ------------------------------------------------------------------ jobs
    def job_exists(self, job_id: str) -> bool:
        return self.store.exists("job", job_id, "")

    def get_job(self, job_id: str) -> dict:
        try:
            return self.store.get("job", job_id, "")
        except NotFound:
            raise BackendError(f"job {job_id} does not exist") from None

    def list_jobs(self, pool_id: Optional[str] = None) -> list[dict]:
        jobs = self.store.query("job")
        return [j for j in jobs if pool_id is None or j.get("pool_id") == pool_id]

    def add_job(self, job: dict) -> dict:
        jid = job.get("id")
        if job.get("schedule_id") and isinstance(jid, str) and jid.startswith(job["schedule_id"] + ":job-"):
            validate_id("job schedule", job["schedule_id"])          # "<schedule id>:job-<n>", generated by the agent
            if not jid[len(job["schedule_id"]) + 5:].isdigit():
                raise BackendError(f"invalid job id {jid!r}")
        else:
            validate_id("job", jid)
        pool = self.get_pool(job["pool_id"])
        from ..utils.versions import check_metadata_compat
        check_metadata_compat(pool.get("metadata", {}))
        rec = {"state": "active", "created": _now(), "state_transition_time": _now(), "priority": 0,
               "max_task_retries": 0, "auto_complete": False, "uses_task_dependencies": False,
               "on_task_failure": "no_action", "env": {}, "job_preparation": None, "job_release": None,
               "prep_nodes": [], "release_done": False, "metadata": {"BATCH_SHIPYARD_VERSION": __version__},
               "terminate_reason": None, "schedule_id": None}
        rec.update(job)
        try:
            self.store.insert("job", rec["id"], "", rec)
        except EntityExists:
            raise BackendError(f"job {rec['id']} already exists") from None
        return rec

    def update_job(self, job_id: str, **patch) -> dict:
        return self.store.merge("job", job_id, "", patch)

    def set_job_state(self, job_id: str, state: str, **extra) -> dict:
        return self.store.merge("job", job_id, "", dict(extra, state=state, state_transition_time=_now()))

    def terminate_job(self, job_id: str, reason: str = "terminated by user") -> None:
        job = self.get_job(job_id)
        if job["state"] in ("completed", "deleting"):
            return
        for t in self.list_tasks(job_id):
            if t["state"] != "completed":
                self.terminate_task(job_id, t["id"], reason="job terminated")
        to_release = bool(job.get("job_release")) and not job.get("release_done")
        if not to_release:
            self.clean_mi_containers(job_id)          # otherwise the agent does it after the job release command
        self.set_job_state(job_id, "terminating" if to_release else "completed", terminate_reason=reason)

    # ------------------------------------------------------------------ named "containers"
    @staticmethod
    def _session_pids(sid: int) -> list[int]:
        out = []
        for d in os.listdir("/proc"):
            if not d.isdigit():
                continue
            try:
                with open(f"/proc/{d}/stat") as f:
                    st = f.read()
                if int(st[st.rindex(")") + 2:].split()[3]) == sid:
                    out.append(int(d))
            except (OSError, ValueError, IndexError):
                continue
        return out

    def clean_mi_containers(self, job_id: str) -> list[str]:
        """Kill what the multi-instance coordination commands of a job left running (the `docker run -d` containers the reference
        removes in job release / `jobs cmi`, /root/reference/convoy/batch.py:2322-2380, 5329-5334).  The runner records the session
        id of every coordination command under <pool>/containers/<name>.coord; returns the names cleaned."""
        try:
            job = self.get_job(job_id)
        except BackendError:
            return []
        cdir = os.path.join(self.pool_root(job["pool_id"]), "containers")
        cleaned = []
        for t in self.list_tasks(job_id):
            if not t.get("multi_instance"):
                continue
            name = re.sub(r"[^A-Za-z0-9_.-]", "_", str((t.get("sandbox") or {}).get("name") or t["id"]))[:128]
            path = os.path.join(cdir, name + ".coord")
            try:
                with open(path) as f:
                    sids = [int(x) for x in f.read().split()]
            except (OSError, ValueError):
                continue
            for sid in sids:
                for sig in (signal.SIGTERM, signal.SIGKILL):
                    pids = [p for p in self._session_pids(sid) if p != os.getpid()]
                    if not pids:
                        break
                    for p in pids:
                        try:
                            os.kill(p, sig)
                        except (ProcessLookupError, PermissionError):
                            pass
                    time.sleep(0.05)
            try:
                os.remove(path)
            except OSError:
                pass
            cleaned.append(name)
        return cleaned

    def delete_job(self, job_id: str) -> None:
        """Terminate, wait for the runners to exit, give the node slots back, then drop rows and files (in that order: the agent
        finalises a task through its row, so rows may only disappear once nothing is running any more)."""
        job = self.get_job(job_id)
        self.set_job_state(job_id, "deleting")
        for t in self.list_tasks(job_id):
            if t["state"] != "completed":
                self.terminate_task(job_id, t["id"], reason="job deleted")
        for t in self.list_tasks(job_id):
            self._reap_task_runner(job, t)
        self.clean_mi_containers(job_id)
        self.store.delete("task", job_id)
        self.store.delete("job", job_id, "")
        shutil.rmtree(os.path.dirname(self.job_dir(job["pool_id"], job_id)), ignore_errors=True)

    def disable_job(self, job_id: str, action: str = "requeue") -> None:
        """action: requeue | terminate | wait — what happens to running tasks."""
        if action not in ("requeue", "terminate", "wait"):
            raise BackendError(f"bad disable action {action}")
        self.get_job(job_id)
        if action in ("requeue", "terminate"):
            for t in self.list_tasks(job_id):
                if t["state"] in ("running", "preparing"):
                    self.terminate_task(job_id, t["id"], reason="job disabled", requeue=(action == "requeue"))
        self.set_job_state(job_id, "disabled", disable_action=action)

    def enable_job(self, job_id: str) -> None:
        job = self.get_job(job_id)
        if job["state"] not in ("disabled", "disabling"):
            raise BackendError(f"job {job_id} is not disabled (state={job['state']})")
        self.set_job_state(job_id, "active")

    def migrate_job(self, job_id: str, pool_id: str) -> None:
        """Re-target a *disabled* job at another pool."""
        job = self.get_job(job_id)
        if job["state"] != "disabled":
            raise BackendError("a job must be disabled before it can be migrated")
        self.get_pool(pool_id)
        old = job["pool_id"]
        self.update_job(job_id, pool_id=pool_id, prep_nodes=[])
        src, dst = os.path.dirname(self.job_dir(old, job_id)), os.path.dirname(self.job_dir(pool_id, job_id))
        if os.path.isdir(src) and src != dst:
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.move(src, dst)

    # ------------------------------------------------------------------ tasks
    def add_tasks(self, job_id: str, tasks: Iterable[dict]) -> int:
        """Submit in chunks of at most 100 (the service's collection limit, kept as a knob)."""
        self.get_job(job_id)
        tasks = list(tasks)
        added = 0
        for i in range(0, len(tasks), MAX_TASKS_PER_REQUEST):
            rows, seen = [], set()
            for t in tasks[i:i + MAX_TASKS_PER_REQUEST]:
                rec = {"state": "active", "created": _now(), "state_transition_time": _now(), "retry_count": 0,
                       "exit_code": None, "result": None, "node_ids": [], "start_time": None, "end_time": None,
                       "pid": None, "failure_info": None, "requeue_count": 0}
                rec.update(t)
                validate_id("task", rec.get("id"))
                if rec["id"] in seen:
                    raise BackendError(f"task {rec['id']} already exists in job {job_id}")
                seen.add(rec["id"])
                rows.append((rec["id"], rec))
            try:
                added += self.store.insert_many("task", job_id, rows)      # one transaction per collection of <= 100 tasks
            except EntityExists as e:
                raise BackendError(f"task {str(e).rsplit('/', 1)[-1]} already exists in job {job_id}") from None
        return added

    def get_task(self, job_id: str, task_id: str) -> dict:
        try:
            return self.store.get("task", job_id, task_id)
        except NotFound:
            raise BackendError(f"task {task_id} does not exist in job {job_id}") from None

    def list_tasks(self, job_id: str) -> list[dict]:
        return self.store.query("task", job_id)

    def task_ids(self, job_id: str) -> set:
        return {t["id"] for t in self.list_tasks(job_id)}

    def update_task(self, job_id: str, task_id: str, **patch) -> dict:
        if "state" in patch:
            patch["state_transition_time"] = _now()
        return self.store.merge("task", job_id, task_id, patch)

    def count_tasks(self, job_id: str) -> dict:
        c = {"active": 0, "running": 0, "completed": 0, "succeeded": 0, "failed": 0}
        for t in self.list_tasks(job_id):
            st = "running" if t["state"] in ("running", "preparing") else t["state"]
            c[st] = c.get(st, 0) + 1
            if t["state"] == "completed":
                c["succeeded" if t.get("result") == "success" else "failed"] += 1
        return c

    def terminate_task(self, job_id: str, task_id: str, reason: str = "terminated", requeue: bool = False,
                       force: bool = False) -> None:
        """SIGTERM to the runner, which tears its ranks down (SIGTERM, 5 s, SIGKILL per rank process group), runs the epilogue and
        writes result.json.  The runner itself is never SIGKILLed here: its ranks live in their own process groups and would survive
        it.  ``force`` (the reference's `docker kill` side channel, /root/reference/convoy/batch.py:2722-2742) additionally SIGKILLs
        the rank process groups right away instead of waiting for the runner's 5 s grace period."""
        t = self.get_task(job_id, task_id)
        if t["state"] == "completed":
            return
        pid = t.get("pid")
        if t["state"] in ("running", "preparing") and pid:
            if pid_alive(pid, t.get("pid_start")):
                try:
                    os.kill(int(pid), signal.SIGTERM)
                except (ProcessLookupError, PermissionError):
                    pass
                if force:
                    self._kill_rank_groups(self.get_job(job_id), t)
            self.update_task(job_id, task_id, terminate_requested=True, terminate_reason=reason, requeue_on_exit=requeue)
            return
        if requeue:
            self.update_task(job_id, task_id, state="active")
        else:
            self.update_task(job_id, task_id, state="completed", result="failure", exit_code=None, end_time=_now(),
                             failure_info={"category": "usererror", "code": "TaskEnded", "message": reason})

    def release_task_slots(self, pool_id: str, job_id: str, task_id: str, ok: bool = False) -> int:
        """Remove [job, task] from every node of the pool that lists it (idempotent); returns the number of slots released."""
        released = 0
        for n in self.store.query("node", pool_id):
            if not any(list(x) == [job_id, task_id] for x in n.get("running_tasks") or []):
                continue

            def fn(x):
                x["running_tasks"] = [y for y in x["running_tasks"] if list(y) != [job_id, task_id]]
                x["total_tasks_run"] = int(x.get("total_tasks_run") or 0) + 1
                x["total_tasks_succeeded"] = int(x.get("total_tasks_succeeded") or 0) + (1 if ok else 0)
                if x["state"] == "running" and not x["running_tasks"]:
                    x["state"] = "idle"; x["state_transition_time"] = _now()
                elif x["state"] == "leaving_pool" and not x["running_tasks"]:
                    x["state"] = "offline"
            try:
                m = self.store.mutate("node", pool_id, n["id"], fn)
                if m["state"] == "offline":
                    self.store.delete("node", pool_id, n["id"])
                released += 1
            except NotFound:
                pass
        return released

    def _kill_rank_groups(self, job: dict, t: dict) -> None:
        """SIGKILL every rank process group the runner recorded in <taskdir>/ranks.pid."""
        tdir = self.task_dir(job["pool_id"], job["id"], t["id"])
        try:
            with open(os.path.join(tdir, "ranks.pid")) as f:
                pgids = [int(x) for x in f.read().split()]
        except (OSError, ValueError):
            pgids = []
        for g in pgids:
            if g > 1:
                try:
                    os.killpg(g, signal.SIGKILL)
                except (ProcessLookupError, PermissionError):
                    pass

    def _reap_task_runner(self, job: dict, t: dict, grace: float = 12.0) -> None:
        """Wait for a terminated task's runner to exit; escalate to the rank process groups, then the runner; free its slots."""
        pid, start = t.get("pid"), t.get("pid_start")
        if t["state"] in ("running", "preparing") and pid:
            deadline = time.time() + grace          # the runner escalates to SIGKILL on its ranks after 5 s, then runs the epilogue
            while pid_alive(pid, start) and time.time() < deadline:
                time.sleep(0.02)
            if pid_alive(pid, start):
                self._kill_rank_groups(job, t)
                try:
                    os.kill(int(pid), signal.SIGKILL)
                except (ProcessLookupError, PermissionError):
                    pass
        self.release_task_slots(job["pool_id"], job["id"], t["id"], ok=False)

    def delete_task(self, job_id: str, task_id: str) -> None:
        t = self.get_task(job_id, task_id)
        job = self.get_job(job_id)
        self.terminate_task(job_id, task_id, reason="task deleted")
        self._reap_task_runner(job, self.get_task(job_id, task_id))
        self.store.delete("task", job_id, task_id)
        shutil.rmtree(self.task_dir(job["pool_id"], job_id, t["id"]), ignore_errors=True)

    def job_stats(self, job_id: Optional[str] = None) -> dict:
        jobs = [self.get_job(job_id)] if job_id else self.list_jobs()
        tot = {"jobs": len(jobs), "tasks": 0, "active": 0, "running": 0, "completed": 0, "succeeded": 0, "failed": 0,
               "retries": 0, "wall_time_s": 0.0, "wait_time_s": 0.0}
        durations, e2e, walltime, job_spans = [], [], [], []
        now = _now()
        for j in jobs:
            if j.get("state") == "completed" and j.get("end_time") or j.get("state") == "completed" and j.get("state_transition_time"):
                job_spans.append((j.get("end_time") or j["state_transition_time"]) - j["created"])
            for t in self.list_tasks(j["id"]):
                tot["tasks"] += 1
                st = "running" if t["state"] in ("running", "preparing") else t["state"]
                tot[st] += 1
                tot["retries"] += int(t.get("retry_count") or 0)
                if t["state"] == "completed":
                    tot["succeeded" if t.get("result") == "success" else "failed"] += 1
                    if t.get("start_time") and t.get("end_time"):
                        d = t["end_time"] - t["start_time"]
                        durations.append(d); tot["wall_time_s"] += d
                        walltime.append(d)
                        e2e.append(t["end_time"] - t["created"])
                        tot["wait_time_s"] += max(0.0, t["start_time"] - t["created"])
                elif st == "running" and t.get("start_time"):
                    walltime.append(now - t["start_time"])

        def agg(v):
            return {"min": min(v, default=None), "max": max(v, default=None), "mean": sum(v) / len(v) if v else None}

        tot["task_wall_time_s"] = agg(durations)
        # the three timing blocks `jobs stats` of the reference prints (/root/reference/convoy/batch.py:2040-2099)
        tot["job_creation_to_completion_s"] = agg(job_spans)
        tot["task_end_to_end_s"] = agg(e2e)                       # creation -> end, completed tasks
        tot["task_command_walltime_s"] = agg(walltime)            # start -> end (or now), running and completed tasks
        tot["completed_pct_of_total"] = round(100.0 * tot["completed"] / tot["tasks"], 2) if tot["tasks"] else None
        tot["succeeded_pct_of_completed"] = round(100.0 * tot["succeeded"] / tot["completed"], 2) if tot["completed"] else None
        tot["failed_pct_of_completed"] = round(100.0 * tot["failed"] / tot["completed"], 2) if tot["completed"] else None
        return tot

    # ------------------------------------------------------------------ files
    def list_task_files(self, job_id: str, task_id: str) -> list[dict]:
        job = self.get_job(job_id)
        base = self.task_dir(job["pool_id"], job_id, task_id)
        out = []
        for d, _, fs in os.walk(base):
            for fn in fs:
                p = os.path.join(d, fn)
                st = os.stat(p)
                out.append({"name": os.path.relpath(p, base).replace(os.sep, "/"), "size": st.st_size,
                            "modified": _iso(st.st_mtime)})
        return sorted(out, key=lambda x: x["name"])

    def task_file_path(self, job_id: str, task_id: str, name: str) -> str:
        job = self.get_job(job_id)
        base = self.task_dir(job["pool_id"], job_id, task_id)
        p = os.path.normpath(os.path.join(base, name))
        if not p.startswith(base + os.sep):
            raise BackendError("file path escapes the task directory")
        return p

    # ------------------------------------------------------------------ job schedules
    def add_job_schedule(self, sched: dict) -> dict:
        validate_id("job schedule", sched.get("id"))
        rec = {"state": "active", "created": _now(), "last_run": None, "runs": 0, "active_job_id": None}
        rec.update(sched)
        try:
            self.store.insert("jobschedule", rec["id"], "", rec)
        except EntityExists:
            raise BackendError(f"job schedule {rec['id']} already exists") from None
        return rec

    def list_job_schedules(self) -> list[dict]:
        return self.store.query("jobschedule")

    def get_job_schedule(self, sid: str) -> dict:
        try:
            return self.store.get("jobschedule", sid, "")
        except NotFound:
            raise BackendError(f"job schedule {sid} does not exist") from None

    def terminate_job_schedule(self, sid: str) -> None:
        s = self.get_job_schedule(sid)
        self.store.merge("jobschedule", sid, "", {"state": "completed"})
        if s.get("active_job_id") and self.job_exists(s["active_job_id"]):
            self.terminate_job(s["active_job_id"], reason="job schedule terminated")

    def delete_job_schedule(self, sid: str) -> None:
        self.terminate_job_schedule(sid)
        self.store.delete("jobschedule", sid, "")
