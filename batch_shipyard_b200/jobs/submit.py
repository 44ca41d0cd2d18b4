"""``jobs add``: turn jobs.yaml into job + task objects on the local backend.

Mirrors the flow of /root/reference/convoy/batch.py:5056-5897 (``add_jobs``):
auto-pool check, per-job pass 1 (task-factory expansion, missing-image policy,
dependency / exit-condition detection), job preparation (image wait + job
input_data + user command) and release tasks, recurrence -> job schedule with a
stored task map, existing-job compatibility checks (tasks are appended),
per-job auto-scratch multi-instance task (block | dependency), pass 2 (task
construction with generated ids), merge task depending on every task of the
submission, chunked task-collection submission, auto_complete, federation
hand-off.
"""
from __future__ import annotations

import datetime
import shlex
import time
from typing import Optional

from ..backend.local import BackendError, LocalBackend, validate_id
from ..config import settings as S
from ..utils import util
from . import builder as B

logger = util.get_logger()
MAX_MERGE_DEP_CHARS = 64000


class JobSubmissionError(RuntimeError):
    pass


def check_jobs_for_auto_pool(config: dict) -> bool:
    """auto_pool must be set on all jobs or none."""
    specs = S.job_specifications(config)
    flags = [S.job_auto_pool(j) is not None for j in specs]
    if any(flags) and not all(flags):
        raise JobSubmissionError("auto_pool must be specified for all jobs or for none of them")
    return bool(flags) and all(flags)


def _parse_ts(v: Optional[str]) -> Optional[float]:
    if not v:
        return None
    try:
        return datetime.datetime.fromisoformat(str(v).replace("Z", "+00:00")).timestamp()
    except ValueError:
        raise JobSubmissionError(f"cannot parse timestamp '{v}' (ISO-8601 expected)") from None


def _auto_scratch_task(pool: S.PoolSettings, jobspec: dict, counts: dict) -> Optional[dict]:
    a = S.job_auto_scratch(jobspec)
    if a is None:
        return None
    if not pool.per_job_auto_scratch:
        raise JobSubmissionError("job auto_scratch needs per_job_auto_scratch: true on the pool")
    n = S.resolve_num_instances(a.num_instances, counts["current_dedicated"], counts["current_low_priority"],
                                pool.vm_dedicated, pool.vm_low_priority)
    jid = S.job_id(jobspec)
    scratch = f"$AZ_BATCH_NODE_SHARED_DIR/auto_scratch/{shlex.quote(jid)}"
    return {"id": a.task_id, "job_id": jid, "command": ":", "image": None, "runtime": "process", "native_shape": False,
            "multi_instance": {"num_instances": n, "coordination_command": f"mkdir -p {scratch} && chmod 1777 {scratch}",
                               "processes_per_node": 1, "mpi": None, "pre_execution_command": None, "resource_files": []},
            "env": {}, "env_exclude": list(B.ENV_EXCLUDE), "depends_on": [], "max_task_retries": 0,
            "exit_job_action": "none", "exit_dependency_action": "block", "is_auto_scratch": True}


def add_jobs(b: LocalBackend, config: dict, recreate: bool = False, tail: Optional[str] = None,
             federation_id: Optional[str] = None, dry_run: bool = False, pool_override: Optional[str] = None) -> dict:
    """Submit every job of ``job_specifications``.  Returns ``{job_id: {...}}`` (the --raw shape)."""
    from ..pool import topology
    autopool = check_jobs_for_auto_pool(config)
    pool = S.pool_settings(config)
    if pool_override:
        pool.id = pool_override
    gs = S.global_settings(config)
    out: dict = {}
    for jobspec in S.job_specifications(config):
        jid = S.job_id(jobspec)
        try:
            validate_id("job schedule" if S.job_recurrence(jobspec) is not None else "job", jid)
        except BackendError as e:          # ids become directory names and shell words (the Batch service enforced this for the reference)
            if not dry_run:
                raise JobSubmissionError(str(e)) from None
            logger.warning("dry run: %s (a real submission is rejected)", e)     # templates carry placeholders such as '<job id>'
        recurrence = S.job_recurrence(jobspec)
        auto_scratch = S.job_auto_scratch(jobspec)
        if recurrence is not None and auto_scratch is not None:
            raise JobSubmissionError("auto_scratch is incompatible with recurrence")
        if federation_id and auto_scratch is not None:
            raise JobSubmissionError("auto_scratch is incompatible with federations")
        if recurrence is not None and not (S.job_auto_complete(jobspec) or recurrence.jm_monitor_task_completion):
            raise JobSubmissionError("a recurring job needs auto_complete: true or job_manager.monitor_task_completion")
        pool_id = pool.id
        if autopool:
            if len(pool.id) > 20:
                raise JobSubmissionError("pool id must be at most 20 characters for auto_pool")
            pool_id = f"{pool.id}-{jid}"[:64]
        if not dry_run and federation_id is None:
            if autopool and not b.pool_exists(pool_id):
                from ..pool.provision import create_pool
                import copy
                cfg = copy.deepcopy(config)
                cfg["pool_specification"]["id"] = pool_id
                create_pool(b, cfg)
                b.store.merge("pool", pool_id, "", {"auto_pool": S.job_auto_pool(jobspec), "auto_pool_job": jid})
            if not b.pool_exists(pool_id):
                raise JobSubmissionError(f"pool {pool_id} does not exist; run `shipyard pool add` first")
        counts = b.current_node_counts(pool_id) if (not dry_run and b.pool_exists(pool_id)) else \
            {"current_dedicated": pool.vm_dedicated, "current_low_priority": pool.vm_low_priority}
        gpu_count = len((b.get_pool(pool_id).get("gpus") or [])) if (not dry_run and b.pool_exists(pool_id)) else topology.gpu_count()
        if gpu_count == 0:
            gpu_count = S.local_gpu_count_from_vm_size(pool.vm_size) or 0

        # ---- pass 1: expand + policy checks -----------------------------------------------------
        expanded = list(S.job_tasks(config, jobspec))
        if not expanded:
            raise JobSubmissionError(f"job {jid} has no tasks")
        merge = S.job_merge_task(jobspec)
        uses_deps = S.job_force_enable_task_dependencies(jobspec) or merge is not None or \
            any(t.get("depends_on") or t.get("depends_on_range") for t in expanded) or \
            (auto_scratch is not None and auto_scratch.setup == "dependency")
        has_exit = any(((t.get("exit_conditions") or {}).get("default") or {}).get("exit_options") for t in expanded) or \
            bool(((jobspec.get("exit_conditions") or {}).get("default") or {}).get("exit_options"))

        # ---- existing job: append after compatibility checks --------------------------------------
        existing_ids: set = set()
        job_exists = (not dry_run) and federation_id is None and b.job_exists(jid)
        if job_exists:
            if recreate:
                b.delete_job(jid)
                job_exists = False
            else:
                ej = b.get_job(jid)
                if ej["state"] in ("completed", "terminating", "deleting"):
                    raise JobSubmissionError(f"job {jid} exists in state {ej['state']}; use --recreate")
                from ..utils.versions import check_metadata_compat
                check_metadata_compat(ej.get("metadata") or {}, what=f"job {jid}")
                if uses_deps and not ej.get("uses_task_dependencies"):
                    raise JobSubmissionError(f"existing job {jid} was created without task dependencies; cannot add dependent tasks")
                if has_exit and ej.get("on_task_failure") != "perform_exit_options_job_action":
                    raise JobSubmissionError(f"existing job {jid} was created without exit-condition job actions")
                existing_ids = b.task_ids(jid)

        # ---- job-level pieces -----------------------------------------------------------------------
        prep_cmds = []
        if not pool.native and (gs.docker_images or gs.singularity_images_unsigned or gs.singularity_images_signed):
            prep_cmds.append(f"$SHIPYARD_PYTHON -m batch_shipyard_b200.pool.wait_images --state-dir $SHIPYARD_STATE_DIR --pool {shlex.quote(pool_id)}")
        if S.job_preparation_command(jobspec):
            prep_cmds.append(S.job_preparation_command(jobspec))
        rel_cmds = []
        if auto_scratch is not None:
            rel_cmds.append(f"rm -rf $AZ_BATCH_NODE_SHARED_DIR/auto_scratch/{shlex.quote(jid)}")
        if S.job_release_command(jobspec):
            rel_cmds.append(S.job_release_command(jobspec))
        job_rec = {
            "id": jid, "pool_id": pool_id, "priority": S.job_priority(jobspec),
            "max_task_retries": S.job_max_task_retries(jobspec),
            "max_wall_time_s": None if S.job_max_wall_time(jobspec) is None else S.job_max_wall_time(jobspec).total_seconds(),
            "retention_time_s": S.job_retention_time(jobspec).total_seconds(),
            "auto_complete": S.job_auto_complete(jobspec), "uses_task_dependencies": uses_deps,
            "on_task_failure": "perform_exit_options_job_action" if has_exit else "no_action",
            "env": S.job_environment_variables(jobspec),
            "job_preparation": {"command": "; ".join(prep_cmds)} if (prep_cmds or jobspec.get("input_data")) else None,
            "job_release": {"command": "; ".join(rel_cmds)} if rel_cmds else None,
            "input_data": jobspec.get("input_data"), "user_identity": jobspec.get("user_identity"),
        }
        if job_rec["job_preparation"] and not job_rec["job_preparation"]["command"]:
            job_rec["job_preparation"]["command"] = ":"

        # ---- pass 2: construct tasks -------------------------------------------------------------------
        reserved: set = set()
        records: list[dict] = []
        scratch = _auto_scratch_task(pool, jobspec, counts) if auto_scratch is not None else None
        if scratch is not None:
            records.append(scratch)
            reserved.add(scratch["id"])
        ids = B.TaskIdAllocator(existing_ids, reserved)       # one scan of the job's ids, then a counter per prefix
        for t in expanded:
            prefix, zfill = t.pop("##autoid")
            tid = t.get("id")
            if not tid:
                tid = ids.next(prefix, zfill)
            elif tid in existing_ids or tid in reserved:
                raise JobSubmissionError(f"task id {tid} already exists in job {jid}")
            else:
                ids.reserve(tid)
            reserved.add(tid)
            try:
                rec = B.build_task(config, pool, jobspec, t, tid, counts, gpu_count=gpu_count, dry_run=dry_run).to_dict()
            except ValueError as e:
                raise JobSubmissionError(str(e)) from e
            if scratch is not None and auto_scratch.setup == "dependency":
                rec["depends_on"] = list(rec["depends_on"]) + [scratch["id"]]
            if scratch is not None:
                rec["env"].setdefault("SHIPYARD_AUTO_SCRATCH", f"$AZ_BATCH_NODE_SHARED_DIR/auto_scratch/{jid}")
            records.append(rec)
        if merge is not None:
            gid = S.global_settings(config).autogenerated_task_id
            jauto = S.autogenerated_task_id(jobspec.get("autogenerated_task_id"), gid)
            mid = merge.get("id") or ids.next(jauto.prefix, jauto.zfill_width, is_merge=True)
            mrec = B.build_task(config, pool, jobspec, merge, mid, counts, is_merge=True, gpu_count=gpu_count, dry_run=dry_run).to_dict()
            mrec["depends_on"] = [r["id"] for r in records if not r.get("is_auto_scratch")]
            if sum(len(x) + 1 for x in mrec["depends_on"]) > MAX_MERGE_DEP_CHARS:
                raise JobSubmissionError("merge task depends on too many tasks (dependency id list exceeds 64000 characters)")
            records.append(mrec)

        summary = {"pool_id": pool_id, "num_tasks": len(records), "task_ids": [r["id"] for r in records],
                   "uses_task_dependencies": uses_deps}
        if dry_run:
            summary["dry_run"] = True
            summary["tasks"] = [{"id": r["id"], "container_command": r.get("container_command"), "mpi_command": r.get("mpi_command"),
                                 "multi_instance": r.get("multi_instance"), "depends_on": r.get("depends_on")} for r in records]
            out[jid] = summary
            continue

        # ---- federation hand-off: the job is not created here ---------------------------------------------
        if federation_id is not None:
            from ..fed import client as fedclient
            info = fedclient.submit_job_to_federation(b, config, federation_id, job_rec, records, jobspec,
                                                      recurrence=recurrence)
            out[jid] = info
            continue

        # ---- recurrence: job schedule + stored task map -------------------------------------------------------
        if recurrence is not None:
            sched = {"id": jid, "pool_id": pool_id, "recurrence_interval_s": recurrence.interval.total_seconds(),
                     "do_not_run_until_ts": _parse_ts(recurrence.do_not_run_until),
                     "do_not_run_after_ts": _parse_ts(recurrence.do_not_run_after),
                     "start_window_s": None if recurrence.start_window is None else recurrence.start_window.total_seconds(),
                     "job_manager": {"allow_low_priority_node": recurrence.jm_allow_low_priority,
                                     "run_exclusive": recurrence.jm_run_exclusive,
                                     "monitor_task_completion": recurrence.jm_monitor_task_completion},
                     "job_template": {k: v for k, v in job_rec.items() if k != "id"}, "task_map": records}
            if b.store.exists("jobschedule", jid, ""):
                if not recreate:
                    raise JobSubmissionError(f"job schedule {jid} already exists; use --recreate")
                b.delete_job_schedule(jid)
            b.add_job_schedule(sched)
            summary.update({"kind": "job_schedule", "tasks_per_recurrence": len(records)})
            out[jid] = summary
            continue

        if not job_exists:
            b.add_job(job_rec)
        n = b.add_tasks(jid, records)
        summary["tasks_added"] = n
        if scratch is not None and auto_scratch.setup == "block":
            summary["auto_scratch"] = "block"
        out[jid] = summary
    return out


def wait_for_tasks(b: LocalBackend, job_id: str, timeout: Optional[float] = None, poll: float = 0.2) -> dict:
    t0 = time.time()
    while True:
        c = b.count_tasks(job_id)
        if c["active"] == 0 and c["running"] == 0:
            return c
        if timeout is not None and time.time() - t0 > timeout:
            return c
        time.sleep(poll)
