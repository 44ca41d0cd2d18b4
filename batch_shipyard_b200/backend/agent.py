"""Node agent: schedules tasks onto the pool's GPUs and drives the native runner.

The part of Azure Batch the reference never had to write: picking runnable tasks
(dependencies, job state, priority), placing them on node slots (pack / spread,
multi-instance tasks need N nodes at once), job preparation / release tasks,
retries, exit-condition actions, auto-complete, job schedules (recurrence, the
reference's /root/reference/cargo/recurrent_job_manager.py:56-253), autoscale evaluation and
recovery of tasks orphaned by an agent crash.  Execution itself is delegated to
``shipyard-taskrun`` (one process per task), so the agent only polls.

One agent per pool holds a TTL lease in the store; a second agent for the same
pool refuses to start (same pattern as the reference's blob leases).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import threading
import time
import uuid
from typing import Optional

from ..state.store import NotFound
from .local import BackendError, LocalBackend, _now, pid_alive, proc_start_ticks
from . import runspec

LEASE_S = 15.0


class NodeAgent:
    def __init__(self, backend: LocalBackend, pool_id: str, poll: float = 0.05, holder: Optional[str] = None):
        self.b, self.pool_id, self.poll = backend, pool_id, poll
        self.holder = holder or f"agent-{os.getpid()}-{uuid.uuid4().hex[:6]}"
        self.procs: dict = {}          # (job, task) -> Popen
        self.lease_name = f"agent-{pool_id}"
        self._last_autoscale = 0.0
        self._last_sample = 0.0
        self.metrics = None
        self.lease_lost = False

    # ------------------------------------------------------------------ lease
    def acquire(self) -> bool:
        return self.b.store.acquire_lease(self.lease_name, self.holder, LEASE_S)

    def release(self) -> None:
        self.b.store.release_lease(self.lease_name, self.holder)

    # ------------------------------------------------------------------ main loop
    def _lease_keeper(self, stop: threading.Event) -> None:
        """Renew the pool lease from a thread of its own: job preparation / release commands run synchronously inside the scheduling
        loop and may take longer than the lease; losing it silently would let a second agent schedule the same tasks."""
        from ..state.store import Store
        st = Store(self.b.store.root)          # sqlite connections are per thread
        while not stop.wait(LEASE_S / 3.0):
            try:
                ok = st.renew_lease(self.lease_name, self.holder, LEASE_S) or st.acquire_lease(self.lease_name, self.holder, LEASE_S)
            except Exception:  # noqa: BLE001 - a locked database is retried on the next beat
                continue
            if not ok:
                self.lease_lost = True
                return

    def _ensure_lease(self) -> bool:
        """Renew (or re-acquire after an expiry nobody used); False means another agent owns the pool now."""
        if self.lease_lost:
            return False
        st = self.b.store
        if st.renew_lease(self.lease_name, self.holder, LEASE_S) or st.acquire_lease(self.lease_name, self.holder, LEASE_S):
            return True
        self.lease_lost = True
        return False

    def run(self, until_idle: bool = True, idle_timeout: float = 0.0, max_seconds: Optional[float] = None) -> None:
        if not self.acquire():
            raise BackendError(f"another agent holds pool {self.pool_id}")
        t0, idle_since = time.time(), None
        stop = threading.Event()
        keeper = threading.Thread(target=self._lease_keeper, args=(stop,), daemon=True, name="lease-keeper")
        keeper.start()
        try:
            self.recover_orphans()
            while True:
                if not self._ensure_lease():
                    # single-agent guarantee lost (lease expired and somebody else took it): stop scheduling.  The runners this agent
                    # started keep running; the new owner adopts them through recover_orphans().
                    raise BackendError(f"agent {self.holder} lost the lease of pool {self.pool_id}; another agent took over")
                progressed = self.tick()
                busy = bool(self.procs) or progressed
                if busy:
                    idle_since = None
                elif until_idle:
                    if not self.has_pending_work():
                        idle_since = idle_since or time.time()
                        if time.time() - idle_since >= idle_timeout:
                            return
                    else:
                        idle_since = None
                if max_seconds is not None and time.time() - t0 > max_seconds:
                    return
                time.sleep(self.poll if busy else max(self.poll, 0.1))
        finally:
            stop.set()
            keeper.join(timeout=2.0)
            self.release()

    def has_pending_work(self) -> bool:
        for job in self.b.list_jobs(self.pool_id):
            if job["state"] == "terminating":
                return True
            if job["state"] != "active":
                continue
            tasks = self.b.list_tasks(job["id"])
            done = {t["id"]: t for t in tasks if t["state"] == "completed"}
            for t in tasks:
                if t["state"] in ("running", "preparing"):
                    return True
                if t["state"] == "active" and self._deps_state(job, t, done, {x["id"] for x in tasks}) != "blocked":
                    return True
        for s in self.b.list_job_schedules():
            if s["state"] == "active" and s.get("pool_id") == self.pool_id:
                return True
        return False

    # ------------------------------------------------------------------ one scheduling pass
    def tick(self) -> bool:
        try:
            self.b.get_pool(self.pool_id)
        except BackendError:
            return False
        progressed = self._reap()
        progressed |= self._sweep_stale_slots()
        progressed |= self._run_schedules()
        progressed |= self._finish_jobs()
        progressed |= self._schedule()
        self._autoscale()
        return progressed

    # -- completion handling ---------------------------------------------------------
    def _reap(self) -> bool:
        progressed = False
        for key, p in list(self.procs.items()):
            rc = p.poll()
            if rc is None:
                continue
            del self.procs[key]
            self._complete(key[0], key[1], None if rc == _AdoptedProc.EXIT_UNKNOWN else rc)
            progressed = True
        return progressed

    def _sweep_stale_slots(self) -> bool:
        """Node slots whose task row no longer exists (deleted while running) or is not running any more and has no runner here."""
        swept = False
        for n in self.b.list_nodes(self.pool_id):
            for jt in list(n.get("running_tasks") or []):
                j, t = jt[0], jt[1]
                if (j, t) in self.procs:
                    continue
                row = self.b.store.try_get("task", j, t)
                if row is None or row.get("state") not in ("running", "preparing"):
                    swept |= self.b.release_task_slots(self.pool_id, j, t, ok=False) > 0
        return swept

    def _complete(self, job_id: str, task_id: str, rc: Optional[int]) -> None:
        """`rc` is the runner's wait status, or None when it is unknown (adopted runner that disappeared)."""
        try:
            t = self.b.get_task(job_id, task_id)
            job = self.b.get_job(job_id)
        except BackendError:
            # the task or job row was deleted while the runner was alive: the slots must still come back
            self.b.release_task_slots(self.pool_id, job_id, task_id, ok=False)
            return
        tdir = self.b.task_dir(job["pool_id"], job_id, task_id)
        result = None
        try:
            with open(os.path.join(tdir, "result.json")) as f:
                result = json.load(f)
        except Exception:  # noqa: BLE001 - runner died before writing the result
            result = None
        if result is None and rc is None and not t.get("terminate_requested"):
            # an adopted runner vanished without a result (SIGKILL, OOM): its exit status is unknowable, so the task did NOT
            # succeed.  Same treatment as recover_orphans(): free the slots and run it again.
            for nid in t.get("node_ids") or []:
                self._release_slot(nid, job_id, task_id, False)
            self.b.update_task(job_id, task_id, state="active", pid=None, pid_start=None, node_ids=[],
                               requeue_count=int(t.get("requeue_count") or 0) + 1, last_exit_code=None)
            return
        exit_code = result["exit_code"] if result else (rc if rc is not None else -1)
        for nid in t.get("node_ids") or []:
            self._release_slot(nid, job_id, task_id, exit_code == 0)
        if t.get("requeue_on_exit"):
            self.b.update_task(job_id, task_id, state="active", pid=None, node_ids=[], requeue_on_exit=False,
                               terminate_requested=False, requeue_count=int(t.get("requeue_count") or 0) + 1)
            return
        max_retries = int(t.get("max_task_retries") or 0)
        retries = int(t.get("retry_count") or 0)
        if exit_code != 0 and not t.get("terminate_requested") and (max_retries < 0 or retries < max_retries):
            self.b.update_task(job_id, task_id, state="active", pid=None, node_ids=[], retry_count=retries + 1,
                               last_exit_code=exit_code)
            return
        ok = exit_code == 0
        info = None
        if not ok:
            cat = "usererror"
            msg = "task terminated" if t.get("terminate_requested") else (
                "wall time exceeded" if result and result.get("timed_out") else f"exit code {exit_code}")
            info = {"category": cat, "code": "FailureExitCode", "message": msg,
                    "rank_exit_codes": (result or {}).get("rank_exit_codes")}
        self.b.update_task(job_id, task_id, state="completed", result="success" if ok else "failure",
                           exit_code=exit_code, end_time=_now(), pid=None, failure_info=info)
        if not ok and not t.get("terminate_requested"):
            action = t.get("exit_job_action") or "none"
            if action == "terminate":
                self.b.terminate_job(job_id, reason=f"task {task_id} failed (exit_conditions job_action: terminate)")
            elif action == "disable":
                self.b.disable_job(job_id, "requeue")

    def _release_slot(self, node_id: str, job_id: str, task_id: str, ok: bool) -> None:
        def fn(n):
            n["running_tasks"] = [x for x in n["running_tasks"] if x != [job_id, task_id] and tuple(x) != (job_id, task_id)]
            n["total_tasks_run"] += 1
            n["total_tasks_succeeded"] += 1 if ok else 0
            if n["state"] == "running" and not n["running_tasks"]:
                n["state"] = "idle"; n["state_transition_time"] = _now()
            elif n["state"] == "leaving_pool" and not n["running_tasks"]:
                n["state"] = "offline"
        try:
            n = self.b.store.mutate("node", self.pool_id, node_id, fn)
            if n["state"] == "offline":
                self.b.store.delete("node", self.pool_id, node_id)
        except NotFound:
            pass

    # -- job lifecycle ---------------------------------------------------------------
    def _finish_jobs(self) -> bool:
        progressed = False
        for job in self.b.list_jobs(self.pool_id):
            jid = job["id"]
            if job["state"] == "active" and job.get("auto_complete"):
                tasks = self.b.list_tasks(jid)
                if tasks and all(t["state"] == "completed" for t in tasks):
                    if not job.get("job_release"):
                        self.b.clean_mi_containers(jid)
                    self.b.set_job_state(jid, "terminating" if job.get("job_release") else "completed",
                                         terminate_reason="AllTasksComplete")
                    progressed = True
                    job = self.b.get_job(jid)
            if job["state"] == "terminating":
                if any(k[0] == jid for k in self.procs):
                    continue          # running tasks are being torn down
                if job.get("job_release") and not job.get("release_done"):
                    for nid in job.get("prep_nodes") or []:
                        self._run_aux(job, "jobrelease", job["job_release"]["command"], nid)
                    self.b.update_job(jid, release_done=True)
                self.b.clean_mi_containers(jid)        # job release also removes the daemonised coordination "containers"
                self.b.set_job_state(jid, "completed")
                progressed = True
        return progressed

    def _run_aux(self, job: dict, kind: str, command: str, node_id: str) -> int:
        """Job preparation / release: short synchronous runner invocation on one node."""
        pool = self.b.get_pool(self.pool_id)
        try:
            node = self.b.get_node(self.pool_id, node_id)
        except BackendError:
            node = {"id": node_id, "gpu_index": None, "dedicated": True}
        spec, tdir = runspec.build_aux_spec(self.b, pool, job, kind, command, node)
        self._ensure_lease()
        rc = subprocess.call([runspec.runner_path(), "--spec", spec], cwd=tdir)      # the lease keeper thread renews meanwhile
        self._ensure_lease()
        return rc

    # -- dependencies ------------------------------------------------------------------
    def _deps_state(self, job: dict, t: dict, done: dict, all_ids: set) -> str:
        """'ready' | 'waiting' | 'blocked' (a dependency failed and the action is block)."""
        deps = list(t.get("depends_on") or [])
        dr = t.get("depends_on_range")
        if dr:
            deps += [str(i) for i in range(int(dr[0]), int(dr[1]) + 1)]
        state = "ready"
        for d in deps:
            if d not in all_ids:
                return "waiting"            # dependency not submitted yet
            dt = done.get(d)
            if dt is None:
                state = "waiting"
                continue
            if dt.get("result") != "success" and (dt.get("exit_dependency_action") or "block") != "satisfy":
                return "blocked"
        return state

    # -- placement ------------------------------------------------------------------------
    def _free_nodes(self, pool: dict, all_nodes: Optional[list] = None) -> list[dict]:
        nodes = [n for n in (self.b.list_nodes(self.pool_id) if all_nodes is None else all_nodes)
                 if n["state"] in ("idle", "running") and n.get("scheduling", "enabled") == "enabled"
                 and len(n["running_tasks"]) < pool["max_tasks_per_node"]]
        if pool.get("node_fill_type", "pack") == "pack":
            nodes.sort(key=lambda n: (-len(n["running_tasks"]), n["ordinal"]))
        else:
            nodes.sort(key=lambda n: (len(n["running_tasks"]), n["ordinal"]))
        return nodes

    def _schedule(self) -> bool:
        pool = self.b.get_pool(self.pool_id)
        progressed = False
        jobs = [j for j in self.b.list_jobs(self.pool_id) if j["state"] == "active"]
        jobs.sort(key=lambda j: (-int(j.get("priority") or 0), j["created"]))
        # the node table is read once per pass and again after every launch (the only thing that changes it here): a saturated pool
        # with thousands of queued tasks costs one query per pass instead of two per queued task
        snap: dict = {"nodes": None, "free": None}

        def nodes_now():
            if snap["nodes"] is None:
                snap["nodes"] = self.b.list_nodes(self.pool_id)
                snap["free"] = self._free_nodes(pool, snap["nodes"])
            return snap["nodes"], snap["free"]

        for job in jobs:
            tasks = self.b.list_tasks(job["id"])
            if not tasks:
                continue
            done = {t["id"]: t for t in tasks if t["state"] == "completed"}
            all_ids = {t["id"] for t in tasks}
            for t in sorted(tasks, key=lambda x: x["created"]):
                if t["state"] != "active":
                    continue
                if self._deps_state(job, t, done, all_ids) != "ready":
                    continue
                mi = t.get("multi_instance")
                need = int(mi["num_instances"]) if mi else 1
                all_nodes, free = nodes_now()
                if mi:
                    free = [n for n in free if not n["running_tasks"]] if need > 1 else free
                if len(free) < need:
                    total = len([n for n in all_nodes if n["state"] in ("idle", "running")])
                    if mi and need > total and total > 0 and not t.get("_warned"):
                        self.b.update_task(job["id"], t["id"], _warned=True, scheduling_note=(
                            f"needs {need} instances but pool has {total} usable node(s)"))
                    continue
                chosen = free[:need]
                if self._launch(pool, job, t, chosen):
                    progressed = True
                snap["nodes"] = None                      # running_tasks / states changed
        return progressed

    def _launch(self, pool: dict, job: dict, t: dict, nodes: list[dict]) -> bool:
        jid, tid = job["id"], t["id"]
        # job preparation runs once per (job, node) before the first task of the job on that node
        if job.get("job_preparation"):
            prepped = set(job.get("prep_nodes") or [])
            for n in nodes:
                if n["id"] not in prepped:
                    rc = self._run_aux(job, "jobpreparation", job["job_preparation"]["command"], n["id"])
                    if rc != 0:
                        self.b.update_task(jid, tid, state="completed", result="failure", exit_code=rc, end_time=_now(),
                                           failure_info={"category": "usererror", "code": "JobPreparationFailed",
                                                         "message": f"job preparation exited with {rc} on {n['id']}"})
                        return True
                    prepped.add(n["id"])
            if sorted(prepped) != sorted(job.get("prep_nodes") or []):
                self.b.update_job(jid, prep_nodes=sorted(prepped))
                job["prep_nodes"] = sorted(prepped)       # the caller's snapshot is reused for the other tasks of this scheduling pass
        spec, tdir = runspec.build_task_spec(self.b, pool, job, t, nodes)
        for n in nodes:
            def fn(x, _j=jid, _t=tid):
                x["running_tasks"].append([_j, _t])
                x["state"] = "running"; x["state_transition_time"] = _now()
            self.b.store.mutate("node", self.pool_id, n["id"], fn)
        env = dict(os.environ)
        p = subprocess.Popen([runspec.runner_path(), "--spec", spec], cwd=tdir, env=env,
                             stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, start_new_session=True)
        self.procs[(jid, tid)] = p
        self.b.update_task(jid, tid, state="running", pid=p.pid, pid_start=proc_start_ticks(p.pid),
                           node_ids=[n["id"] for n in nodes], start_time=_now(), agent=self.holder)
        return True

    # -- crash recovery ---------------------------------------------------------------------
    def recover_orphans(self) -> int:
        """Tasks marked running by a dead agent: adopt their result if the runner finished, else requeue."""
        n = 0
        for job in self.b.list_jobs(self.pool_id):
            for t in self.b.list_tasks(job["id"]):
                if t["state"] not in ("running", "preparing") or (job["id"], t["id"]) in self.procs:
                    continue
                pid = t.get("pid")
                if pid and pid_alive(pid, t.get("pid_start")):
                    self.procs[(job["id"], t["id"])] = _AdoptedProc(int(pid), t.get("pid_start"))
                    continue
                tdir = self.b.task_dir(job["pool_id"], job["id"], t["id"])
                if os.path.exists(os.path.join(tdir, "result.json")):
                    self._complete(job["id"], t["id"], None)
                else:
                    for nid in t.get("node_ids") or []:
                        self._release_slot(nid, job["id"], t["id"], False)
                    self.b.update_task(job["id"], t["id"], state="active", pid=None, node_ids=[],
                                       requeue_count=int(t.get("requeue_count") or 0) + 1)
                n += 1
        return n

    # -- job schedules (recurrence) -------------------------------------------------------------
    def _run_schedules(self) -> bool:
        progressed = False
        now = _now()
        for s in self.b.list_job_schedules():
            if s["state"] != "active" or s.get("pool_id") != self.pool_id:
                continue
            if s.get("do_not_run_after_ts") and now > s["do_not_run_after_ts"]:
                self.b.store.merge("jobschedule", s["id"], "", {"state": "completed"})
                continue
            if s.get("do_not_run_until_ts") and now < s["do_not_run_until_ts"]:
                continue
            active = s.get("active_job_id")
            if active and self.b.job_exists(active) and self.b.get_job(active)["state"] not in ("completed",):
                continue          # at most one active job per schedule
            last = s.get("last_run")
            if last is not None and now - last < float(s["recurrence_interval_s"]):
                continue
            run_no = int(s.get("runs") or 0) + 1
            jid = f"{s['id']}:job-{run_no}"
            job = dict(s["job_template"], id=jid, schedule_id=s["id"], pool_id=self.pool_id, auto_complete=True)
            self.b.add_job(job)
            self.b.add_tasks(jid, [dict(t) for t in s["task_map"]])
            self.b.store.merge("jobschedule", s["id"], "", {"last_run": now, "runs": run_no, "active_job_id": jid})
            progressed = True
        return progressed

    # -- autoscale --------------------------------------------------------------------------------
    def _autoscale(self) -> None:
        pool = self.b.get_pool(self.pool_id)
        a = pool.get("autoscale") or {}
        if not a.get("enabled") or not a.get("formula"):
            return
        from ..pool import autoscale as AS
        now = _now()
        if self.metrics is None:
            self.metrics = AS.MetricsWindow(sample_period=float(a.get("sample_period_s") or 5.0))
        if now - self._last_sample >= self.metrics.sample_period:
            self._last_sample = now
            active = running = 0
            for j in self.b.list_jobs(self.pool_id):
                c = self.b.count_tasks(j["id"])
                active += c["active"]; running += c["running"]
            self.metrics.add("$ActiveTasks", now, active)
            self.metrics.add("$RunningTasks", now, running)
            self.metrics.add("$PendingTasks", now, active + running)
            self.metrics.add("$PreemptedNodeCount", now, 0)
        interval = float(a.get("evaluation_interval_s") or 900.0)
        if now - self._last_autoscale < interval:
            return
        self._last_autoscale = now
        self.evaluate_autoscale(apply=True)

    def evaluate_autoscale(self, apply: bool = False) -> dict:
        from ..pool import autoscale as AS
        import datetime as _dt
        pool = self.b.get_pool(self.pool_id)
        a = pool.get("autoscale") or {}
        if not a.get("formula"):
            raise BackendError(f"pool {self.pool_id} has no autoscale formula")
        m = self.metrics or AS.MetricsWindow()
        cur = self.b.current_node_counts(self.pool_id)
        m.current = {"$CurrentDedicatedNodes": cur["current_dedicated"], "$CurrentLowPriorityNodes": cur["current_low_priority"]}
        res = AS.FormulaInterpreter(m, _dt.datetime.now()).run(a["formula"])
        ngpu = len(pool.get("gpus") or [])
        d, lp = res.target_dedicated, res.target_low_priority
        if ngpu:
            d = min(d, ngpu); lp = min(lp, ngpu - d)
        out = {"target_dedicated": d, "target_low_priority": lp, "node_deallocation_option": res.node_deallocation_option,
               "timestamp": _now(), "error": None}
        self.b.store.mutate("pool", self.pool_id, "", lambda p: p["autoscale"].__setitem__("last_evaluation", out))
        if apply and (d != cur["current_dedicated"] or lp != cur["current_low_priority"]) and d + lp > 0:
            self.b.resize_pool(self.pool_id, d, lp)
            from ..pool.provision import bring_up_nodes
            bring_up_nodes(self.b, self.pool_id)
        return out


class _AdoptedProc:
    """poll()-compatible handle for a runner started by a previous agent process.  It is not our child, so its exit status cannot
    be read: poll() returns None while it runs and EXIT_UNKNOWN afterwards; result.json is then the only source of truth."""
    EXIT_UNKNOWN = "unknown"

    def __init__(self, pid: int, start_ticks=None):
        self.pid, self.start_ticks = pid, start_ticks

    def poll(self):
        return None if pid_alive(self.pid, self.start_ticks) else self.EXIT_UNKNOWN


def spawn_detached_agent(state_dir: str, pool_id: str, idle_timeout: float = 20.0) -> Optional[int]:
    """Start a background agent for the pool unless one already holds the lease."""
    from ..state.store import Store
    st = Store(state_dir)
    if st.lease_holder(f"agent-{pool_id}") is not None:
        return None
    logdir = os.path.join(state_dir, "pools", pool_id)
    os.makedirs(logdir, exist_ok=True)
    log = open(os.path.join(logdir, "agent.log"), "ab")
    p = subprocess.Popen([sys.executable, "-m", "batch_shipyard_b200.backend.agent", "--state-dir", state_dir,
                          "--pool", pool_id, "--idle-timeout", str(idle_timeout)],
                         stdout=log, stderr=log, stdin=subprocess.DEVNULL, start_new_session=True,
                         cwd=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    log.close()                               # the agent holds its own copy of the descriptor
    return p.pid


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="shipyard node agent (one per pool)")
    ap.add_argument("--state-dir", required=True)
    ap.add_argument("--pool", required=True)
    ap.add_argument("--idle-timeout", type=float, default=20.0)
    ap.add_argument("--forever", action="store_true")
    a = ap.parse_args(argv)
    b = LocalBackend(state_dir=a.state_dir)
    agent = NodeAgent(b, a.pool)
    try:
        agent.run(until_idle=not a.forever, idle_timeout=a.idle_timeout)
    except BackendError as e:
        print(f"agent: {e}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
