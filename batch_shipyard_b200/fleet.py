"""Orchestration layer: one ``action_*`` per CLI verb.

Counterpart of /root/reference/convoy/fleet.py (99 ``action_*`` functions,
:2974-5447).  Every verb of the reference CLI has an action here; verbs that
only make sense against Azure (subscriptions, ARM disks, key vault, RDP) operate
on the local analogue (the box, the local state store, a local secret store) or
report that they are not applicable on a local pool — they never fail the CLI.

Each action takes a ``Context`` (merged config + backend) and returns a plain
JSON-able value; the CLI prints it (``--raw`` = JSON, otherwise a readable form).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import time
from dataclasses import dataclass, field
from typing import Optional

from . import __version__
from .backend.agent import NodeAgent, spawn_detached_agent
from .backend.local import BackendError, LocalBackend
from .config import settings as S
from .state.store import default_state_dir, entity_names
from .utils import util

logger = util.get_logger()


class ActionError(RuntimeError):
    pass


@dataclass
class Context:
    config: dict
    state_dir: str
    backend: LocalBackend = field(default=None)  # type: ignore[assignment]
    raw: bool = False
    yes: bool = False
    verbose: bool = False
    inline_agent: bool = False     # run the node agent in-process (tests / --wait) instead of detaching

    def __post_init__(self):
        if self.backend is None:
            self.backend = LocalBackend(state_dir=self.state_dir)

    @property
    def b(self) -> LocalBackend:
        return self.backend

    def pool_id(self, override: Optional[str] = None) -> str:
        return override or S.pool_id(self.config)

    def confirm(self, msg: str) -> bool:
        return self.yes or util.confirm_action({"_auto_confirm": self.yes}, msg)


def resolve_state_dir(config: dict) -> str:
    return os.environ.get("SHIPYARD_STATE_DIR") or S.credentials_local_state_dir(config) or default_state_dir()


def _na(what: str, **extra) -> dict:
    return dict({"status": "not_applicable", "detail": f"{what} is a cloud operation; nothing to do on a local B200 pool"}, **extra)


def _kick_agent(ctx: Context, pool_id: str, wait: bool = False, timeout: Optional[float] = None) -> None:
    """Make sure something is executing the pool's tasks."""
    if ctx.inline_agent or wait:
        agent = NodeAgent(ctx.b, pool_id)
        try:
            agent.run(until_idle=True, idle_timeout=0.0, max_seconds=timeout)
        except BackendError:
            # a detached agent already owns the pool: just wait for it
            t0 = time.time()
            while NodeAgent(ctx.b, pool_id).has_pending_work():
                if timeout and time.time() - t0 > timeout:
                    break
                time.sleep(0.2)
        return
    spawn_detached_agent(ctx.state_dir, pool_id)


# =============================================================================== account
def action_account_info(ctx: Context, name=None, resource_group=None) -> dict:
    from .pool import topology
    return {"account": "local", "state_dir": ctx.state_dir, "version": __version__, "topology": topology.describe(),
            "pools": [p["id"] for p in ctx.b.list_pools()]}


def action_account_list(ctx: Context, resource_group=None) -> list:
    return [action_account_info(ctx)]


def action_account_quota(ctx: Context, location=None) -> dict:
    from .pool import topology
    n = topology.gpu_count()
    used = sum(len(p.get("gpus") or []) for p in ctx.b.list_pools())
    return {"location": location or "local", "gpu_quota": n, "gpus_in_pools": used, "pool_quota": None,
            "cpu_slots": os.cpu_count()}


def action_account_images(ctx: Context, show_unrelated=False, show_unverified=False) -> list:
    from .pool.cascade import image_store_dir
    d = image_store_dir(ctx.state_dir)
    out = []
    if os.path.isdir(d):
        for fn in sorted(os.listdir(d)):
            p = os.path.join(d, fn)
            out.append({"artefact": fn, "bytes": os.path.getsize(p) if os.path.isfile(p) else None})
    return out


# =============================================================================== pool
def action_pool_add(ctx: Context, recreate=False, no_wait=False) -> dict:
    from .pool import provision
    try:
        pool = provision.create_pool(ctx.b, ctx.config, recreate=recreate, no_wait=no_wait)
    except (provision.PoolCreationError, ValueError, BackendError) as e:
        raise ActionError(str(e)) from e
    ps = S.pool_settings(ctx.config)
    gs = S.global_settings(ctx.config)
    # pool-level input data and `transfer_files_on_pool_creation` ingress
    if ps.input_data or (ps.transfer_files_on_pool_creation and gs.files):
        from .data import ingress
        if ps.input_data:
            ingress.pool_input_data(ctx.b, ctx.config, ps.id, ps.input_data)
        if ps.transfer_files_on_pool_creation:
            ingress.ingress_data(ctx.b, ctx.config, ps.id, to_fs=None)
    counts = ctx.b.node_counts(ps.id)
    return {ps.id: {"allocation_state": ctx.b.get_pool(ps.id)["allocation_state"], "node_counts": counts,
                    "summary": pool.get("_summary"), "gpus": pool.get("gpus"),
                    "global_resources": [{"resource": e["resource"], "state": e["state"]} for e in ctx.b.store.query("globalresource", ps.id)]}}


def action_pool_exists(ctx: Context, pool_id=None) -> bool:
    return ctx.b.pool_exists(ctx.pool_id(pool_id))


def action_pool_list(ctx: Context) -> dict:
    out = {}
    for p in ctx.b.list_pools():
        c = ctx.b.current_node_counts(p["id"])
        out[p["id"]] = {"state": p["state"], "allocation_state": p["allocation_state"], "vm_size": p["vm_size"],
                        "current_dedicated": c["current_dedicated"], "current_low_priority": c["current_low_priority"],
                        "target_dedicated": p["target_dedicated"], "target_low_priority": p["target_low_priority"],
                        "max_tasks_per_node": p["max_tasks_per_node"], "gpus": p.get("gpus"),
                        "autoscale_enabled": bool((p.get("autoscale") or {}).get("enabled"))}
    return out


def action_pool_delete(ctx: Context, pool_id=None, wait=False) -> dict:
    pid = ctx.pool_id(pool_id)
    if not ctx.b.pool_exists(pid):
        raise ActionError(f"pool {pid} does not exist")
    if not ctx.confirm(f"delete pool {pid}"):
        return {"deleted": False}
    ctx.b.delete_pool(pid)
    return {"deleted": True, "pool_id": pid}


def action_pool_resize(ctx: Context, wait=False) -> dict:
    from .pool import provision
    ps = S.pool_settings(ctx.config)
    try:
        ctx.b.resize_pool(ps.id, ps.vm_dedicated, ps.vm_low_priority)
    except BackendError as e:
        raise ActionError(str(e)) from e
    summary = provision.bring_up_nodes(ctx.b, ps.id, ps)
    return {ps.id: {"node_counts": ctx.b.node_counts(ps.id), "summary": summary}}


def action_pool_stats(ctx: Context, pool_id=None) -> dict:
    try:
        return ctx.b.pool_stats(ctx.pool_id(pool_id))
    except BackendError as e:
        raise ActionError(str(e)) from e


def action_pool_ssh(ctx: Context, cardinal=None, nodeid=None, tty=False, command=()) -> dict:
    """There is no remote host: run the command (or report the node's shell context) locally."""
    pid = ctx.pool_id()
    nodes = ctx.b.list_nodes(pid)
    if not nodes:
        raise ActionError(f"pool {pid} has no nodes")
    node = next((n for n in nodes if n["id"] == nodeid), None) if nodeid else nodes[int(cardinal or 0)]
    if node is None:
        raise ActionError(f"node {nodeid} not found in pool {pid}")
    env = dict(os.environ, AZ_BATCH_POOL_ID=pid, AZ_BATCH_NODE_ID=node["id"], AZ_BATCH_NODE_ROOT_DIR=ctx.b.pool_root(pid),
               AZ_BATCH_NODE_SHARED_DIR=ctx.b.node_shared_dir(pid))
    if node.get("gpu_index") is not None:
        env["CUDA_VISIBLE_DEVICES"] = str(node["gpu_index"])
    if not command:
        return {"node": node["id"], "login": "local", "cwd": ctx.b.pool_root(pid), "env": {k: env[k] for k in env if k.startswith(("AZ_BATCH", "CUDA_VIS"))}}
    p = subprocess.run(" ".join(command), shell=True, env=env, cwd=ctx.b.pool_root(pid), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return {"node": node["id"], "exit_code": p.returncode, "output": p.stdout}


def action_pool_rdp(ctx: Context, cardinal=None, no_auto=False, nodeid=None) -> dict:
    return _na("RDP")


def action_pool_user_add(ctx: Context) -> dict:
    ps = S.pool_settings(ctx.config)
    if not ps.ssh_username:
        raise ActionError("pool ssh.username is not configured")
    from . import crypto
    export = ps.ssh_generated_file_export_path
    priv, pub = crypto.generate_ssh_keypair(export, prefix="id_rsa_shipyard")
    ctx.b.store.insert("pooluser", ps.id, ps.ssh_username, {"public_key": pub, "private_key": priv,
                       "expiry": time.time() + ps.ssh_expiry_days * 86400}, replace=True)
    return {"username": ps.ssh_username, "private_key": priv, "public_key": pub, "note": "local pool: key recorded, no remote account needed"}


def action_pool_user_del(ctx: Context) -> dict:
    ps = S.pool_settings(ctx.config)
    n = ctx.b.store.delete("pooluser", ps.id, ps.ssh_username or "")
    return {"deleted": n}


def action_pool_autoscale_enable(ctx: Context) -> dict:
    from .pool import autoscale as AS
    ps = S.pool_settings(ctx.config)
    if ps.autoscale is None:
        raise ActionError("pool_specification.autoscale is not configured")
    formula = AS.get_formula(ps)
    ctx.b.store.mutate("pool", ps.id, "", lambda p: p["autoscale"].update(
        {"enabled": True, "formula": formula, "evaluation_interval_s": ps.autoscale.evaluation_interval.total_seconds()}))
    return {"pool_id": ps.id, "enabled": True, "formula": formula}


def action_pool_autoscale_disable(ctx: Context) -> dict:
    pid = ctx.pool_id()
    ctx.b.store.mutate("pool", pid, "", lambda p: p["autoscale"].update({"enabled": False}))
    return {"pool_id": pid, "enabled": False}


def action_pool_autoscale_evaluate(ctx: Context) -> dict:
    from .pool import autoscale as AS
    ps = S.pool_settings(ctx.config)
    pool = ctx.b.get_pool(ps.id)
    if not (pool.get("autoscale") or {}).get("formula"):
        if ps.autoscale is None:
            raise ActionError("no autoscale formula to evaluate")
        ctx.b.store.mutate("pool", ps.id, "", lambda p: p["autoscale"].update({"formula": AS.get_formula(ps)}))
    return NodeAgent(ctx.b, ps.id).evaluate_autoscale(apply=False)


def action_pool_autoscale_lastexec(ctx: Context) -> dict:
    pool = ctx.b.get_pool(ctx.pool_id())
    return (pool.get("autoscale") or {}).get("last_evaluation") or {"status": "never evaluated"}


def action_pool_images_list(ctx: Context) -> dict:
    pid = ctx.pool_id()
    return {pid: [{"resource": e["resource"], "state": e["state"], "size": e.get("size"), "seconds": e.get("seconds"),
                   "error": e.get("error")} for e in ctx.b.store.query("globalresource", pid)]}


def action_pool_images_update(ctx: Context, docker_image=None, docker_image_digest=None, singularity_image=None, ssh=False) -> dict:
    """Re-load (or add) images on every node: the reference's multi-instance 'broadcast exec' idiom."""
    from .pool import cascade as C
    pid = ctx.pool_id()
    res = S.global_resources_images(ctx.config)
    if docker_image:
        res = [f"docker:{docker_image}" + (f"@{docker_image_digest}" if docker_image_digest else "")]
    if singularity_image:
        res = [f"singularity:{singularity_image}"]
    have = {e["resource"] for e in ctx.b.store.query("globalresource", pid)}
    import hashlib
    for r in res:
        ctx.b.store.insert("globalresource", pid, hashlib.sha1(r.encode()).hexdigest(),
                           {"resource": r, "state": "pending", "size": None, "seconds": None, "error": None}, replace=True)
    gs = S.global_settings(ctx.config)
    ok = C.Cascade(ctx.b.store, pid, concurrency=gs.concurrent_source_downloads).run(block=True)
    return {"pool_id": pid, "updated": res, "new": [r for r in res if r not in have], "ok": ok}


def action_pool_nodes_list(ctx: Context, start_task_failed=False, unusable=False) -> dict:
    pid = ctx.pool_id()
    return {pid: [{"node_id": n["id"], "state": n["state"], "gpu_index": n.get("gpu_index"), "dedicated": n["dedicated"],
                   "running_tasks": len(n["running_tasks"]), "total_tasks_run": n["total_tasks_run"],
                   "errors": n.get("errors") or []} for n in ctx.b.list_nodes(pid, start_task_failed, unusable)]}


def action_pool_nodes_count(ctx: Context, pool_id=None) -> dict:
    pid = ctx.pool_id(pool_id)
    try:
        return {pid: ctx.b.node_counts(pid)}
    except BackendError as e:
        raise ActionError(str(e)) from e


def action_pool_nodes_grls(ctx: Context, no_generate_tunnel_script=False) -> dict:
    pid = ctx.pool_id()
    return {pid: [{"node_id": n["id"], "ip": "127.0.0.1", "port": 22, "gpu_index": n.get("gpu_index")} for n in ctx.b.list_nodes(pid)]}


def action_pool_nodes_del(ctx: Context, all_start_task_failed=False, all_starting=False, all_unusable=False, nodeid=()) -> dict:
    pid = ctx.pool_id()
    ids = set(nodeid or ())
    for n in ctx.b.list_nodes(pid):
        if (all_start_task_failed and n["state"] == "start_task_failed") or (all_starting and n["state"] == "starting") or \
                (all_unusable and n["state"] == "unusable"):
            ids.add(n["id"])
    if not ids:
        raise ActionError("no nodes selected")
    if not ctx.confirm(f"delete node(s) {sorted(ids)} from pool {pid}"):
        return {"deleted": []}
    for nid in ids:
        ctx.b.remove_node(pid, nid)
    return {"deleted": sorted(ids)}


def action_pool_nodes_reboot(ctx: Context, all_start_task_failed=False, nodeid=()) -> dict:
    from .pool import provision
    pid = ctx.pool_id()
    ids = set(nodeid or ())
    if all_start_task_failed:
        ids |= {n["id"] for n in ctx.b.list_nodes(pid, start_task_failed=True)}
    for nid in ids:
        ctx.b.set_node_state(pid, nid, "rebooting")
        m = os.path.join(ctx.b.node_startup_dir(pid), nid, provision.NODEPREP_FINISHED)
        if os.path.exists(m):
            pass   # marker kept: the reboot takes the fast path, like the reference's node prep
    summary = provision.bring_up_nodes(ctx.b, pid, S.pool_settings(ctx.config) if "pool_specification" in ctx.config else None)
    return {"rebooted": sorted(ids), "summary": summary}


def action_pool_nodes_ps(ctx: Context) -> dict:
    pid = ctx.pool_id()
    out = {}
    for n in ctx.b.list_nodes(pid):
        out[n["id"]] = [{"job_id": j, "task_id": t, "pid": ctx.b.get_task(j, t).get("pid")} for j, t in n["running_tasks"]]
    return out


def action_pool_nodes_zap(ctx: Context, no_remove=False, stop=False) -> dict:
    """Kill every task process on the pool (the reference: `docker kill/rm` all containers)."""
    pid = ctx.pool_id()
    killed = []
    for j in ctx.b.list_jobs(pid):
        for t in ctx.b.list_tasks(j["id"]):
            if t["state"] in ("running", "preparing"):
                ctx.b.terminate_task(j["id"], t["id"], reason="zap", force=not stop)
                killed.append([j["id"], t["id"]])
    return {"zapped": killed}


def action_pool_nodes_prune(ctx: Context, volumes=False) -> dict:
    """Remove data of completed tasks past their retention time (the reference: `docker system prune`)."""
    pid = ctx.pool_id()
    now, freed = time.time(), 0
    for j in ctx.b.list_jobs(pid):
        for t in ctx.b.list_tasks(j["id"]):
            if t["state"] == "completed" and t.get("end_time") and now - t["end_time"] > float(t.get("retention_time_s") or 7 * 86400):
                d = ctx.b.task_dir(pid, j["id"], t["id"])
                if os.path.isdir(d):
                    freed += sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(d) for f in fs)
                    shutil.rmtree(d, ignore_errors=True)
    return {"freed_bytes": freed}


# =============================================================================== jobs
def action_jobs_add(ctx: Context, recreate=False, tail=None, wait=False, dry_run=False) -> dict:
    from .jobs import submit
    try:
        out = submit.add_jobs(ctx.b, ctx.config, recreate=recreate, tail=tail, dry_run=dry_run)
    except (submit.JobSubmissionError, ValueError, BackendError, KeyError) as e:
        raise ActionError(str(e)) from e
    if dry_run:
        return out
    pools = {v["pool_id"] for v in out.values() if isinstance(v, dict) and v.get("pool_id")}
    for pid in pools:
        _kick_agent(ctx, pid, wait=wait or bool(tail))
    if tail:
        jid = next(iter(out))
        tid = out[jid]["task_ids"][-1]
        try:
            with open(ctx.b.task_file_path(jid, tid, tail)) as f:
                out[jid]["tail"] = f.read()
        except OSError as e:
            out[jid]["tail"] = f"<{e}>"
        try:                                   # a failed task: say so and show why (stderr tail), instead of an empty tail
            t = ctx.b.get_task(jid, tid)
            out[jid]["task_state"] = t.get("state")
            if t.get("exit_code") not in (None, 0):
                out[jid]["exit_code"] = t.get("exit_code")
                with open(ctx.b.task_file_path(jid, tid, "stderr.txt")) as f:
                    out[jid]["stderr_tail"] = f.read()[-2000:]
        except (OSError, BackendError, KeyError):
            pass
    return out


def action_jobs_list(ctx: Context, jobid=None, jobscheduleid=None) -> dict:
    jobs = {}
    for j in ctx.b.list_jobs():
        if jobid and j["id"] != jobid:
            continue
        jobs[j["id"]] = {"state": j["state"], "pool_id": j["pool_id"], "priority": j.get("priority", 0),
                         "uses_task_dependencies": j.get("uses_task_dependencies"), "auto_complete": j.get("auto_complete"),
                         "task_counts": ctx.b.count_tasks(j["id"]), "terminate_reason": j.get("terminate_reason")}
    scheds = {s["id"]: {"state": s["state"], "runs": s.get("runs"), "active_job_id": s.get("active_job_id"),
                        "recurrence_interval_s": s.get("recurrence_interval_s")}
              for s in ctx.b.list_job_schedules() if not jobscheduleid or s["id"] == jobscheduleid}
    return {"jobs": jobs, "job_schedules": scheds}


def _job_ids(ctx: Context, jobid=None, all_jobs=False) -> list:
    if all_jobs:
        return [j["id"] for j in ctx.b.list_jobs()]
    if jobid:
        return [jobid]
    return [S.job_id(j) for j in S.job_specifications(ctx.config)]


def action_jobs_term(ctx: Context, all_jobs=False, all_jobschedules=False, jobid=None, jobscheduleid=None, termtasks=False, wait=False) -> dict:
    out = {"terminated": [], "schedules": []}
    if jobscheduleid or all_jobschedules:
        for s in ctx.b.list_job_schedules():
            if all_jobschedules or s["id"] == jobscheduleid:
                ctx.b.terminate_job_schedule(s["id"]); out["schedules"].append(s["id"])
        if jobscheduleid and not jobid and not all_jobs:
            return out
    for jid in _job_ids(ctx, jobid, all_jobs):
        if not ctx.b.job_exists(jid):
            continue
        if not ctx.confirm(f"terminate job {jid}"):
            continue
        ctx.b.terminate_job(jid)
        out["terminated"].append(jid)
        _kick_agent(ctx, ctx.b.get_job(jid)["pool_id"], wait=wait)
    return out


def action_jobs_del(ctx: Context, all_jobs=False, all_jobschedules=False, jobid=None, jobscheduleid=None, termtasks=False, wait=False) -> dict:
    out = {"deleted": [], "schedules": []}
    for s in ctx.b.list_job_schedules():
        if all_jobschedules or (jobscheduleid and s["id"] == jobscheduleid):
            ctx.b.delete_job_schedule(s["id"]); out["schedules"].append(s["id"])
    if jobscheduleid and not jobid and not all_jobs:
        return out
    for jid in _job_ids(ctx, jobid, all_jobs):
        if ctx.b.job_exists(jid) and ctx.confirm(f"delete job {jid}"):
            ctx.b.delete_job(jid); out["deleted"].append(jid)
    return out


def action_jobs_cmi(ctx: Context, delete=False) -> dict:
    """Clean up multi-instance leftovers: stray rank processes and coordination artefacts."""
    cleaned = []
    for jid in _job_ids(ctx):
        if not ctx.b.job_exists(jid):
            continue
        for name in ctx.b.clean_mi_containers(jid):          # kill daemonised coordination sessions by container name
            cleaned.append([jid, "container:" + name])
        for t in ctx.b.list_tasks(jid):
            if t.get("multi_instance") and t["state"] == "completed":
                tdir = ctx.b.task_dir(ctx.b.get_job(jid)["pool_id"], jid, t["id"])
                for fn in (".heartbeat", ".shipyard.envlist"):
                    try:
                        os.remove(os.path.join(tdir, fn))
                    except OSError:
                        pass
                cleaned.append([jid, t["id"]])
    return {"cleaned": cleaned}


def action_jobs_migrate(ctx: Context, jobid=None, jobscheduleid=None, poolid=None, requeue=False, terminate=False, wait=False) -> dict:
    if not poolid:
        poolid = ctx.pool_id()
    moved = []
    for jid in _job_ids(ctx, jobid):
        job = ctx.b.get_job(jid)
        if job["state"] != "disabled":
            ctx.b.disable_job(jid, "terminate" if terminate else ("wait" if wait else "requeue"))
            _kick_agent(ctx, job["pool_id"], wait=True, timeout=30)
        ctx.b.migrate_job(jid, poolid)
        ctx.b.enable_job(jid)
        _kick_agent(ctx, poolid)
        moved.append(jid)
    return {"migrated": moved, "pool_id": poolid}


def action_jobs_disable(ctx: Context, jobid=None, jobscheduleid=None, requeue=False, terminate=False, wait=False) -> dict:
    action = "terminate" if terminate else ("wait" if wait else "requeue")
    done = []
    for jid in _job_ids(ctx, jobid):
        ctx.b.disable_job(jid, action); done.append(jid)
    return {"disabled": done, "action": action}


def action_jobs_enable(ctx: Context, jobid=None, jobscheduleid=None) -> dict:
    done = []
    for jid in _job_ids(ctx, jobid):
        ctx.b.enable_job(jid); done.append(jid)
        _kick_agent(ctx, ctx.b.get_job(jid)["pool_id"])
    return {"enabled": done}


def action_jobs_stats(ctx: Context, jobid=None) -> dict:
    return ctx.b.job_stats(jobid)


def action_jobs_tasks_list(ctx: Context, all_jobs=False, jobid=None, poll_until_tasks_complete=False, taskid=None) -> dict:
    out = {}
    for jid in _job_ids(ctx, jobid, all_jobs):
        if not ctx.b.job_exists(jid):
            continue
        if poll_until_tasks_complete:
            from .jobs.submit import wait_for_tasks
            _kick_agent(ctx, ctx.b.get_job(jid)["pool_id"])
            wait_for_tasks(ctx.b, jid)
        def iso(ts):
            return None if not ts else time.strftime("%Y-%m-%dT%H:%M:%S", time.gmtime(ts)) + ("%.3f" % (ts % 1))[1:] + "Z"

        out[jid] = [{"task_id": t["id"], "state": t["state"], "result": t.get("result"), "exit_code": t.get("exit_code"),
                     "node_ids": t.get("node_ids"), "retry_count": t.get("retry_count"),
                     "start_time": t.get("start_time"), "end_time": t.get("end_time"),
                     # the reference prints creation / start / end timestamps and the duration of every task
                     "created_utc": iso(t.get("created")), "start_time_utc": iso(t.get("start_time")), "end_time_utc": iso(t.get("end_time")),
                     "duration_s": (round(t["end_time"] - t["start_time"], 3) if t.get("start_time") and t.get("end_time") else None),
                     "multi_instance": bool(t.get("multi_instance")), "command": t.get("command"),
                     "failure_info": t.get("failure_info")}
                    for t in ctx.b.list_tasks(jid) if not taskid or t["id"] == taskid]
    return out


def action_jobs_tasks_count(ctx: Context, jobid=None) -> dict:
    return {jid: ctx.b.count_tasks(jid) for jid in _job_ids(ctx, jobid) if ctx.b.job_exists(jid)}


def action_jobs_tasks_term(ctx: Context, force=False, jobid=None, taskid=None, wait=False) -> dict:
    done = []
    for jid in _job_ids(ctx, jobid):
        for t in ctx.b.list_tasks(jid):
            if taskid and t["id"] != taskid:
                continue
            if t["state"] != "completed" and ctx.confirm(f"terminate task {jid}/{t['id']}"):
                ctx.b.terminate_task(jid, t["id"], force=force); done.append([jid, t["id"]])
        if wait:
            _kick_agent(ctx, ctx.b.get_job(jid)["pool_id"], wait=True, timeout=30)
    return {"terminated": done}


def action_jobs_tasks_del(ctx: Context, jobid=None, taskid=None, wait=False) -> dict:
    done = []
    for jid in _job_ids(ctx, jobid):
        for t in ctx.b.list_tasks(jid):
            if taskid and t["id"] != taskid:
                continue
            if ctx.confirm(f"delete task {jid}/{t['id']}"):
                ctx.b.delete_task(jid, t["id"]); done.append([jid, t["id"]])
    return {"deleted": done}


# =============================================================================== data
def action_data_ingress(ctx: Context, to_fs=None) -> dict:
    from .data import ingress
    pid = ctx.pool_id() if "pool_specification" in ctx.config else None
    return ingress.ingress_data(ctx.b, ctx.config, pid, to_fs=to_fs)


def action_data_files_list(ctx: Context, jobid=None, taskid=None) -> dict:
    out = {}
    for jid in _job_ids(ctx, jobid):
        for t in ctx.b.list_tasks(jid):
            if taskid and t["id"] != taskid:
                continue
            out[f"{jid}/{t['id']}"] = ctx.b.list_task_files(jid, t["id"])
    return out


def _filespec(filespec: Optional[str]) -> tuple:
    parts = (filespec or "").split(",")
    if len(parts) != 3:
        raise ActionError("filespec must be <jobid>,<taskid>,<filename>")
    return parts[0], parts[1], parts[2]


def action_data_files_stream(ctx: Context, disk=False, filespec=None, follow=True, out=None) -> dict:
    jid, tid, name = _filespec(filespec)
    path = ctx.b.task_file_path(jid, tid, name)
    out = out or sys.stdout
    pos, total = 0, 0
    sink = open(os.path.basename(name), "wb") if disk else None
    while True:
        if os.path.exists(path):
            with open(path, "rb") as f:
                f.seek(pos)
                chunk = f.read()
                pos += len(chunk); total += len(chunk)
                if chunk:
                    if sink:
                        sink.write(chunk)
                    else:
                        out.write(chunk.decode("utf8", "replace")); out.flush()
        t = ctx.b.get_task(jid, tid)
        if t["state"] == "completed" or not follow:
            if os.path.exists(path) and os.path.getsize(path) > pos:
                continue
            break
        time.sleep(0.2)
    if sink:
        sink.close()
    return {"streamed_bytes": total, "task_state": ctx.b.get_task(jid, tid)["state"]}


def action_data_files_task(ctx: Context, all=False, filespec=None, dest=".") -> dict:
    jid, tid, name = _filespec(filespec if filespec and filespec.count(",") == 2 else (filespec or "") + ",")
    copied = []
    if all:
        for f in ctx.b.list_task_files(jid, tid):
            src = ctx.b.task_file_path(jid, tid, f["name"])
            dst = os.path.join(dest, jid, tid, f["name"])
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copy2(src, dst); copied.append(dst)
    else:
        dst = os.path.join(dest, os.path.basename(name))
        shutil.copy2(ctx.b.task_file_path(jid, tid, name), dst); copied.append(dst)
    return {"copied": copied}


def action_data_files_node(ctx: Context, all=False, filespec=None, dest=".") -> dict:
    parts = (filespec or "").split(",")
    if len(parts) != 2:
        raise ActionError("filespec must be <nodeid>,<filename>")
    pid = ctx.pool_id()
    root = ctx.b.pool_root(pid)
    src = os.path.normpath(os.path.join(root, parts[1]))
    if not src.startswith(root):
        raise ActionError("path escapes the node root")
    dst = os.path.join(dest, os.path.basename(src))
    if os.path.isdir(src):
        shutil.copytree(src, dst, dirs_exist_ok=True)
    else:
        shutil.copy2(src, dst)
    return {"copied": [dst]}


# =============================================================================== diag / misc
def action_diag_logs_upload(ctx: Context, cardinal=None, generate_sas=False, nodeid=None, wait=False) -> dict:
    """Collect agent / node-prep / cascade logs of the pool into the diagnostics container."""
    pid = ctx.pool_id()
    names = entity_names(S.global_settings(ctx.config).storage_entity_prefix if "batch_shipyard" in ctx.config else "shipyard")
    container = names["blob_resourcefiles"] + "-diag"
    n = 0
    for d, _, fs in os.walk(ctx.b.pool_root(pid)):
        for fn in fs:
            if fn.endswith(".log") or fn in ("stdout.txt", "stderr.txt") and "startup" in d:
                p = os.path.join(d, fn)
                ctx.b.store.put_blob_from_file(container, f"{pid}/{os.path.relpath(p, ctx.b.pool_root(pid))}", p)
                n += 1
    return {"container": container, "files": n, "path": os.path.join(ctx.state_dir, "blobs", container)}


def action_misc_tensorboard(ctx: Context, jobid=None, taskid=None, logdir=None, image=None) -> dict:
    jid = jobid or S.job_id(S.job_specifications(ctx.config)[0])
    tasks = ctx.b.list_tasks(jid)
    if not tasks:
        raise ActionError(f"job {jid} has no tasks")
    tid = taskid or tasks[-1]["id"]
    path = os.path.join(ctx.b.task_dir(ctx.b.get_job(jid)["pool_id"], jid, tid), "wd", logdir or "")
    cmd = f"tensorboard --logdir {path} --port 6006"
    return {"logdir": path, "command": cmd, "url": "http://127.0.0.1:6006", "note": "no tunnel needed on a local pool; run the command above"}


def action_misc_mirror_images(ctx: Context) -> dict:
    from .pool.cascade import image_store_dir, artefact_name, find_artefact
    res = S.global_resources_images(ctx.config)
    return {"image_store": image_store_dir(ctx.state_dir),
            "resources": [{"resource": r, "artefact": artefact_name(r), "present": find_artefact(ctx.state_dir, r) is not None} for r in res]}


# =============================================================================== storage
def action_storage_clear(ctx: Context, diagnostics_logs=False, poolid=None) -> dict:
    pid = poolid or (ctx.pool_id() if "pool_specification" in ctx.config else None)
    if pid:
        ctx.b.store.delete("globalresource", pid)
        ctx.b.store.clear_events(pid)
    else:
        ctx.b.store.clear_events()
    return {"cleared": pid or "all", "diagnostics_logs": diagnostics_logs}


def action_storage_del(ctx: Context, clear_tables=False, diagnostics_logs=False, poolid=None) -> dict:
    out = action_storage_clear(ctx, diagnostics_logs, poolid)
    if clear_tables and ctx.confirm("delete ALL shipyard state"):
        ctx.b.store.clear_all()
        out["all_state_deleted"] = True
    return out


def action_storage_sas_create(ctx: Context, storage_account=None, path=None, file=False, create=False, list_=False, read=False, write=False, delete=False) -> dict:
    root = S.credentials_storage_local_path(ctx.config, storage_account or "") or os.path.join(ctx.state_dir, "storage", storage_account or "local")
    return {"sas": None, "url": "file://" + os.path.join(root, (path or "").strip("/")), "note": "local storage needs no SAS token"}


# =============================================================================== keyvault / cert
def action_keyvault_add(ctx: Context, name: str) -> dict:
    from . import keyvault
    creds = {"credentials": ctx.config.get("credentials") or {}}
    return keyvault.store_credentials(ctx.state_dir, name, creds)


def action_keyvault_del(ctx: Context, name: str) -> dict:
    from . import keyvault
    return keyvault.delete_secret(ctx.state_dir, name)


def action_keyvault_list(ctx: Context) -> list:
    from . import keyvault
    return keyvault.list_secrets(ctx.state_dir)


def action_cert_create(ctx: Context, file_prefix=None, pfx_password=None) -> dict:
    from . import crypto
    return crypto.generate_pem_pfx_certificates(file_prefix or "shipyard-cert", pfx_password)


def action_cert_add(ctx: Context, file=None, pem_no_certs=False, pem_public_key=False, pfx_password=None) -> dict:
    from . import crypto
    if not file:
        raise ActionError("--file is required")
    thumb = crypto.get_sha1_thumbprint(file, pfx_password)
    with open(file, "rb") as f:
        ctx.b.store.put_blob("certs", thumb, f.read())
    ctx.b.store.insert("cert", "local", thumb, {"file": os.path.basename(file), "added": time.time()}, replace=True)
    return {"sha1_thumbprint": thumb}


def action_cert_list(ctx: Context) -> list:
    return [{"sha1_thumbprint": c["_rk"], "file": c.get("file")} for c in ctx.b.store.query("cert", "local")]


def action_cert_del(ctx: Context, sha1=()) -> dict:
    n = 0
    for t in sha1 or [c["_rk"] for c in ctx.b.store.query("cert", "local")]:
        n += ctx.b.store.delete("cert", "local", t)
        ctx.b.store.delete_blob("certs", t)
    return {"deleted": n}


# =============================================================================== fs (remote fs -> local shared dirs)
def action_fs_disks_add(ctx: Context) -> dict:
    from .fs import remotefs
    return remotefs.create_disks(ctx.b, ctx.config)


def action_fs_disks_del(ctx: Context, all=False, delete_resource_group=False, name=None, resource_group=None, wait=False) -> dict:
    from .fs import remotefs
    return remotefs.delete_disks(ctx.b, ctx.config, name=name, all=all)


def action_fs_disks_list(ctx: Context, resource_group=None, restrict_scope=False) -> list:
    from .fs import remotefs
    return remotefs.list_disks(ctx.b)


def action_fs_cluster_add(ctx: Context, storage_cluster_id: str) -> dict:
    from .fs import remotefs
    return remotefs.create_cluster(ctx.b, ctx.config, storage_cluster_id)


def action_fs_cluster_orchestrate(ctx: Context, storage_cluster_id: str) -> dict:
    from .fs import remotefs
    d = remotefs.create_disks(ctx.b, ctx.config)
    c = remotefs.create_cluster(ctx.b, ctx.config, storage_cluster_id)
    return {"disks": d, "cluster": c}


def action_fs_cluster_resize(ctx: Context, storage_cluster_id: str) -> dict:
    from .fs import remotefs
    return remotefs.resize_cluster(ctx.b, ctx.config, storage_cluster_id)


def action_fs_cluster_expand(ctx: Context, storage_cluster_id: str, no_rebalance=False) -> dict:
    from .fs import remotefs
    return remotefs.expand_cluster(ctx.b, ctx.config, storage_cluster_id, rebalance=not no_rebalance)


def action_fs_cluster_del(ctx: Context, storage_cluster_id: str, **kw) -> dict:
    from .fs import remotefs
    if not ctx.confirm(f"delete storage cluster {storage_cluster_id}"):
        return {"deleted": False}
    return remotefs.delete_cluster(ctx.b, storage_cluster_id, delete_data=bool(kw.get("delete_data_disks")))


def action_fs_cluster_suspend(ctx: Context, storage_cluster_id: str, no_wait=False) -> dict:
    from .fs import remotefs
    return remotefs.set_cluster_state(ctx.b, storage_cluster_id, "suspended")


def action_fs_cluster_start(ctx: Context, storage_cluster_id: str, no_wait=False) -> dict:
    from .fs import remotefs
    return remotefs.set_cluster_state(ctx.b, storage_cluster_id, "running")


def action_fs_cluster_status(ctx: Context, storage_cluster_id: str, detail=False, hosts=False) -> dict:
    from .fs import remotefs
    return remotefs.cluster_status(ctx.b, storage_cluster_id, detail=detail)


def action_fs_cluster_ssh(ctx: Context, storage_cluster_id: str, cardinal=None, hostname=None, tty=False, command=()) -> dict:
    from .fs import remotefs
    st = remotefs.cluster_status(ctx.b, storage_cluster_id)
    if not command:
        return {"login": "local", "cwd": st["path"]}
    p = subprocess.run(" ".join(command), shell=True, cwd=st["path"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return {"exit_code": p.returncode, "output": p.stdout}


# =============================================================================== monitor
def action_monitor_create(ctx: Context) -> dict:
    from .monitor import service
    return service.create(ctx.b, ctx.config)


def action_monitor_add(ctx: Context, poolid=(), remote_fs=()) -> dict:
    from .monitor import service
    return service.add_targets(ctx.b, list(poolid or []) or ([ctx.pool_id()] if "pool_specification" in ctx.config and not remote_fs else []), list(remote_fs or []))


def action_monitor_list(ctx: Context) -> dict:
    from .monitor import service
    return service.list_targets(ctx.b)


def action_monitor_remove(ctx: Context, all=False, poolid=(), remote_fs=()) -> dict:
    from .monitor import service
    return service.remove_targets(ctx.b, all, list(poolid or []), list(remote_fs or []))


def action_monitor_ssh(ctx: Context, tty=False, command=()) -> dict:
    from .monitor import service
    return service.shell(ctx.b, command)


def action_monitor_suspend(ctx: Context, no_wait=False) -> dict:
    from .monitor import service
    return service.stop(ctx.b)


def action_monitor_start(ctx: Context, no_wait=False) -> dict:
    from .monitor import service
    return service.start(ctx.b, ctx.config)


def action_monitor_status(ctx: Context) -> dict:
    from .monitor import service
    return service.status(ctx.b)


def action_monitor_destroy(ctx: Context, **kw) -> dict:
    from .monitor import service
    if not ctx.confirm("destroy the monitoring service"):
        return {"destroyed": False}
    return service.destroy(ctx.b)


# =============================================================================== federation
def action_fed_proxy_create(ctx: Context) -> dict:
    from .fed import client
    return client.proxy_create(ctx.b, ctx.config, ctx.state_dir)


def action_fed_proxy_ssh(ctx: Context, tty=False, command=()) -> dict:
    from .fed import client
    return client.proxy_status(ctx.b)


def action_fed_proxy_suspend(ctx: Context, no_wait=False) -> dict:
    from .fed import client
    return client.proxy_stop(ctx.b)


def action_fed_proxy_start(ctx: Context, no_wait=False) -> dict:
    from .fed import client
    return client.proxy_create(ctx.b, ctx.config, ctx.state_dir)


def action_fed_proxy_status(ctx: Context) -> dict:
    from .fed import client
    return client.proxy_status(ctx.b)


def action_fed_proxy_destroy(ctx: Context, **kw) -> dict:
    from .fed import client
    return client.proxy_stop(ctx.b, destroy=True)


def action_fed_create(ctx: Context, federation_id: str, force=False, no_unique_job_ids=False) -> dict:
    from .fed import client
    return client.create_federation(ctx.b, federation_id, force, not no_unique_job_ids)


def action_fed_list(ctx: Context, federation_id=()) -> dict:
    from .fed import client
    return client.list_federations(ctx.b, list(federation_id or []))


def action_fed_destroy(ctx: Context, federation_id: str) -> dict:
    from .fed import client
    if not ctx.confirm(f"destroy federation {federation_id}"):
        return {"destroyed": False}
    return client.destroy_federation(ctx.b, federation_id)


def action_fed_pool_add(ctx: Context, federation_id: str, batch_service_url=None, pool_id=()) -> dict:
    from .fed import client
    pools = list(pool_id or []) or [ctx.pool_id()]
    return client.add_pools(ctx.b, federation_id, pools)


def action_fed_pool_remove(ctx: Context, federation_id: str, all=False, batch_service_url=None, pool_id=()) -> dict:
    from .fed import client
    return client.remove_pools(ctx.b, federation_id, list(pool_id or []), all)


def action_fed_jobs_add(ctx: Context, federation_id: str) -> dict:
    from .jobs import submit
    try:
        return submit.add_jobs(ctx.b, ctx.config, federation_id=federation_id)
    except (submit.JobSubmissionError, ValueError, BackendError) as e:
        raise ActionError(str(e)) from e


def action_fed_jobs_list(ctx: Context, federation_id: str, blocked=False, job_id=None, jobschedule_id=None, queued=False) -> dict:
    from .fed import client
    return client.list_jobs(ctx.b, federation_id, blocked=blocked, queued=queued, job_id=job_id or jobschedule_id)


def action_fed_jobs_term(ctx: Context, federation_id: str, all_jobs=False, all_jobschedules=False, force=False, job_id=(), job_schedule_id=()) -> dict:
    from .fed import client
    return client.enqueue_job_action(ctx.b, federation_id, "terminate", list(job_id or []) + list(job_schedule_id or []), all_jobs or all_jobschedules)


def action_fed_jobs_del(ctx: Context, federation_id: str, all_jobs=False, all_jobschedules=False, job_id=(), job_schedule_id=()) -> dict:
    from .fed import client
    return client.enqueue_job_action(ctx.b, federation_id, "delete", list(job_id or []) + list(job_schedule_id or []), all_jobs or all_jobschedules)


def action_fed_jobs_zap(ctx: Context, federation_id: str, unique_id: str) -> dict:
    from .fed import client
    return client.zap_action(ctx.b, federation_id, unique_id)


# =============================================================================== slurm
def action_slurm_cluster_create(ctx: Context) -> dict:
    from .slurm import cluster
    return cluster.create(ctx.b, ctx.config)


def action_slurm_cluster_orchestrate(ctx: Context, storage_cluster_id=None) -> dict:
    from .slurm import cluster
    out = {}
    if storage_cluster_id:
        out["fs"] = action_fs_cluster_orchestrate(ctx, storage_cluster_id)
    out["slurm"] = cluster.create(ctx.b, ctx.config)
    return out


def action_slurm_cluster_suspend(ctx: Context, no_controller_nodes=False, no_login_nodes=False, no_wait=False) -> dict:
    from .slurm import cluster
    return cluster.set_state(ctx.b, ctx.config, "suspended")


def action_slurm_cluster_start(ctx: Context, no_controller_nodes=False, no_login_nodes=False, no_wait=False) -> dict:
    from .slurm import cluster
    return cluster.set_state(ctx.b, ctx.config, "running")


def action_slurm_cluster_status(ctx: Context) -> dict:
    from .slurm import cluster
    return cluster.status(ctx.b, ctx.config)


def action_slurm_cluster_destroy(ctx: Context, **kw) -> dict:
    from .slurm import cluster
    if not ctx.confirm("destroy the slurm cluster mapping"):
        return {"destroyed": False}
    return cluster.destroy(ctx.b, ctx.config)


def action_slurm_ssh(ctx: Context, kind: str, offset=None, node_name=None, tty=False, command=()) -> dict:
    from .slurm import cluster
    return cluster.shell(ctx.b, ctx.config, kind, node_name, command)
