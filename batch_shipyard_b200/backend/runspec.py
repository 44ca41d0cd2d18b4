"""Build ``shipyard-taskrun`` spec files: the per-task environment contract.

Emulates the Azure Batch environment tasks rely on (all 17 ``AZ_BATCH_*``
variables the reference's recipes/scripts use) plus the
``SHIPYARD_*`` runner variables (/root/reference/convoy/batch.py:4653-4814):
system prologue (resource files, input_data ingress, registry "logins"), user
command, system epilogue (conditional output_data egress), env-file exclusion,
multi-instance coordination / pre-execution commands, wall-time limit.
"""
from __future__ import annotations

import json
import os
import re
import shlex
import sys
import zlib

from .._build import native_dir

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def runner_path() -> str:
    override = os.environ.get("SHIPYARD_TASKRUN_BIN")          # e.g. the Address/UB-sanitizer build used by the tests
    if override and os.path.exists(override):
        return override
    p = os.path.join(native_dir(), "shipyard-taskrun")
    if not os.path.exists(p):
        from .._build import ensure_built
        ensure_built(["taskrun"])
    if not os.path.exists(p):
        raise RuntimeError("shipyard-taskrun is not built (python native/build.py taskrun)")
    return p


def _esc(v: str) -> str:
    return str(v).replace("\\", "\\\\").replace("\n", "\\n").replace("\t", "\\t")


def _write_spec(path: str, items: list[tuple[str, str]]) -> None:
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        for k, v in items:
            f.write(f"{k}\t{_esc(v)}\n")
    os.replace(tmp, path)


def base_env(b, pool: dict, job: dict, task_id: str, tdir: str, nodes: list[dict]) -> dict:
    pid = pool["id"]
    root = b.pool_root(pid)
    master = nodes[0] if nodes else {"id": "cpu-0", "gpu_index": None, "dedicated": True}
    hosts = ",".join("127.0.0.1" for _ in nodes) or "127.0.0.1"
    gpu_ids = [str(n["gpu_index"]) for n in nodes if n.get("gpu_index") is not None]
    env = {
        "AZ_BATCH_ACCOUNT_NAME": "local", "AZ_BATCH_ACCOUNT_URL": "file://" + b.root,
        "AZ_BATCH_AUTHENTICATION_TOKEN": "local", "AZ_BATCH_CERTIFICATES_DIR": os.path.join(root, "certs"),
        "AZ_BATCH_POOL_ID": pid, "AZ_BATCH_JOB_ID": job["id"], "AZ_BATCH_TASK_ID": task_id,
        "AZ_BATCH_NODE_ID": master["id"], "AZ_BATCH_NODE_IS_DEDICATED": "true" if master.get("dedicated", True) else "false",
        "AZ_BATCH_NODE_ROOT_DIR": root, "AZ_BATCH_NODE_SHARED_DIR": b.node_shared_dir(pid),
        "AZ_BATCH_NODE_STARTUP_DIR": b.node_startup_dir(pid), "AZ_BATCH_TASK_DIR": tdir,
        "AZ_BATCH_TASK_WORKING_DIR": os.path.join(tdir, "wd"), "AZ_BATCH_JOB_PREP_DIR": os.path.join(os.path.dirname(tdir), "jobpreparation"),
        "AZ_BATCH_HOST_LIST": hosts, "AZ_BATCH_MASTER_NODE": "127.0.0.1:6000",
        "AZ_BATCH_IS_CURRENT_NODE_MASTER": "true", "AZ_BATCH_NODE_LIST": ";".join(n["id"] for n in nodes),
        "SHIPYARD_POOL_ID": pid, "SHIPYARD_STATE_DIR": b.root, "SHIPYARD_GPUS": ",".join(gpu_ids),
        "SHIPYARD_NUM_GPUS": str(len(gpu_ids)), "SHIPYARD_NUM_INSTANCES": str(max(1, len(nodes))),
        "SHIPYARD_PYTHON": sys.executable, "SHIPYARD_HOME": _REPO_ROOT,
        "SHIPYARD_STAGE_MANIFEST": os.path.join(tdir, ".shipyard.stage.json"),
        "PYTHONPATH": _REPO_ROOT + (os.pathsep + os.environ["PYTHONPATH"] if os.environ.get("PYTHONPATH") else ""),
    }
    if gpu_ids and len(nodes) == 1:
        # a single-instance task sees only the GPU(s) of the node it was placed on
        env["CUDA_VISIBLE_DEVICES"] = ",".join(gpu_ids)
    return env


def _expand(v: str, env: dict) -> str:
    from ..utils.util import expand_env
    return expand_env(v, dict(os.environ, **env))


def _mover(args: list[str]) -> str:
    return " ".join([shlex.quote(sys.executable), "-m", "batch_shipyard_b200.data.mover"] + [shlex.quote(a) for a in args])


def data_commands(b, t: dict, env: dict) -> tuple[list[str], list[str]]:
    """(prologue ingress commands, epilogue egress commands) for resource_files / input_data / output_data."""
    pro, epi = [], []
    wd = env["AZ_BATCH_TASK_WORKING_DIR"]
    for rf in t.get("resource_files") or []:
        src = rf.get("blob_source")
        if not src:
            continue
        args = ["fetch", "--url", src, "--dest", os.path.join(wd, rf["file_path"])]
        if rf.get("file_mode"):
            args += ["--mode", str(rf["file_mode"])]
        pro.append(_mover(args))
    ind = t.get("input_data") or {}
    for spec in ind.get("azure_storage") or []:
        local = _expand(spec.get("local_path") or "$AZ_BATCH_TASK_WORKING_DIR", env)
        args = ["ingress", "--state-dir", b.root, "--link", spec["storage_account_settings"], "--remote", spec["remote_path"],
                "--local", local]
        for inc in spec.get("include") or []:
            args += ["--include", inc]
        for exc in spec.get("exclude") or []:
            args += ["--exclude", exc]
        pro.append(_mover(args))
    for spec in ind.get("azure_batch") or []:
        dest = _expand(spec.get("destination") or "$AZ_BATCH_TASK_WORKING_DIR", env)
        args = ["taskfiles", "--state-dir", b.root, "--job", spec["job_id"], "--task", spec["task_id"], "--dest", dest]
        for inc in spec.get("include") or []:
            args += ["--include", inc]
        for exc in spec.get("exclude") or []:
            args += ["--exclude", exc]
        pro.append(_mover(args))
    for spec in (t.get("output_data") or {}).get("azure_storage") or []:
        local = _expand(spec.get("local_path") or "$AZ_BATCH_TASK_DIR", env)
        args = ["egress", "--state-dir", b.root, "--link", spec["storage_account_settings"], "--remote", spec["remote_path"],
                "--local", local, "--condition", spec.get("condition") or "tasksuccess"]
        for inc in spec.get("include") or []:
            args += ["--include", inc]
        for exc in spec.get("exclude") or []:
            args += ["--exclude", exc]
        epi.append(_mover(args))
    return pro, epi


def _size_bytes(v) -> int:
    """docker --shm-size notation: <number>[b|k|m|g] (default bytes)."""
    m = re.match(r"^\s*(\d+(?:\.\d+)?)\s*([bkmgBKMG]?)[bB]?\s*$", str(v))
    if not m:
        return 0
    return int(float(m.group(1)) * {"": 1, "b": 1, "k": 1 << 10, "m": 1 << 20, "g": 1 << 30}[m.group(2).lower()])


def containers_dir(b, pool_id: str) -> str:
    """Registry of named "containers" (running tasks, daemonised multi-instance coordination sessions) of a pool."""
    return os.path.join(b.pool_root(pool_id), "containers")


def sandbox_items(b, pool: dict, t: dict, tdir: str, env: dict) -> list[tuple[str, str]]:
    """Spec keys that make ``shipyard-taskrun`` run the task in a process sandbox with the container semantics of the job:
    volume binds, restrict_default_bind_mounts, user_identity, --shm-size, --rm, --name
    (/root/reference/convoy/settings.py:3875-3901, 3919-4051).  ``SHIPYARD_SANDBOX`` = off | auto | require overrides the mode."""
    sbx = t.get("sandbox")
    mode = os.environ.get("SHIPYARD_SANDBOX_MODE", "auto")
    if not sbx or mode == "off":
        return []
    scratch = os.path.join(tdir, ".container")
    # The task's own temporary directory is always handed over as TMPDIR.  Mounting it OVER /tmp (what a container has) is opt-in
    # (SHIPYARD_SANDBOX_PRIVATE_TMP=1): tasks here are host programs, and interpreters, checkouts or datasets that live under
    # /tmp would vanish from their view.  The node root and SHIPYARD_HOME are re-attached when they live under /tmp.
    private = os.environ.get("SHIPYARD_SANDBOX_PRIVATE_TMP", "0") not in ("0", "", "off", "false")
    env["TMPDIR"] = os.path.join(scratch, "tmp")
    os.makedirs(env["TMPDIR"], exist_ok=True)
    items = [("sandbox", mode), ("container_scratch", scratch)] + ([("private_tmp", os.path.join(scratch, "tmp"))] if private else []) + [
             ("rm", "1" if sbx.get("remove_after_exit", True) else "0"),
             ("name", re.sub(r"[^A-Za-z0-9_.-]", "_", str(sbx.get("name") or t["id"]))[:128]),
             ("containers_dir", containers_dir(b, pool["id"])), ("node_root", b.pool_root(pool["id"])), ("keep_tmp", _REPO_ROOT)]
    for src, dst, opts, vname in sbx.get("binds") or []:
        src = _expand(src, env) if src else os.path.join(scratch, "volumes", re.sub(r"[^A-Za-z0-9_.-]", "_", vname))
        dst = _expand(dst, env)
        os.makedirs(src, exist_ok=True)
        items.append(("bind", f"{src}:{dst}" + (":ro" if "ro" in str(opts).split(",") else "")))
    if sbx.get("restrict_default_bind_mounts"):
        items += [("restrict_root", b.pool_root(pool["id"])), ("keep", tdir)]
    if sbx.get("uid") is not None:
        items += [("uid", str(int(sbx["uid"]))), ("gid", str(int(sbx.get("gid") if sbx.get("gid") is not None else sbx["uid"])))]
    if sbx.get("shm_size"):
        items.append(("shm_bytes", str(_size_bytes(sbx["shm_size"]))))
    return items


def build_task_spec(b, pool: dict, job: dict, t: dict, nodes: list[dict]) -> tuple[str, str]:
    """Write the runner spec for task `t` placed on `nodes`; returns (spec path, task dir)."""
    pid, jid, tid = pool["id"], job["id"], t["id"]
    tdir = b.task_dir(pid, jid, tid)
    wd = os.path.join(tdir, "wd")
    os.makedirs(wd, exist_ok=True)
    for stale in ("result.json",):
        try:
            os.remove(os.path.join(tdir, stale))
        except FileNotFoundError:
            pass
    env = base_env(b, pool, job, tid, tdir, nodes)
    user_env = dict(job.get("env") or {})
    user_env.update(t.get("env") or {})
    for k, v in user_env.items():
        env[k] = _expand(str(v), env)
    mi = t.get("multi_instance")
    items: list[tuple[str, str]] = [("workdir", wd), ("taskdir", tdir), ("stdout", os.path.join(tdir, "stdout.txt")),
                                    ("stderr", os.path.join(tdir, "stderr.txt")), ("result_file", os.path.join(tdir, "result.json")),
                                    ("heartbeat", os.path.join(tdir, ".heartbeat")),
                                    ("env_file", os.path.join(tdir, ".shipyard.envlist"))]
    for name in t.get("env_exclude") or []:
        items.append(("env_exclude", name))
    pro, epi = data_commands(b, t, env)
    pro = list(t.get("system_prologue") or []) + pro
    epi = epi + list(t.get("system_epilogue") or [])
    if pro:
        items.append(("system_prologue", "set -e; " + "; ".join(pro)))
    if epi:
        items.append(("system_epilogue", "; ".join(epi)))
    command = t.get("command") or ":"
    # the rendezvous name is global to the machine (abstract UNIX socket + POSIX shm): two shipyard installations with different state
    # directories may run pools / jobs / tasks of the same names at the same time, so the state directory is part of it
    inst = zlib.crc32(os.path.realpath(b.root).encode()) & 0xffffff
    session = f"{pid}-{jid}-{tid}-{t.get('retry_count', 0)}-{t.get('requeue_count', 0)}-{inst:06x}"
    if len(session) > 72:
        # the socket name is cut at 106 characters by the transport (sockaddr_un) and programs append their own suffix: long ids are
        # folded into a digest so that the distinguishing tail (attempt counters, installation) can never be truncated away
        import hashlib
        session = session[:40] + "-" + hashlib.sha1(session.encode()).hexdigest()[:28]
    # a stable, collision-resistant rendezvous port per task attempt
    port = 20000 + (zlib.crc32(session.encode()) % 20000)
    items += [("session", session), ("master_port", str(port))]
    preload_shim = False
    if mi:
        n, ppn = int(mi["num_instances"]), int(mi.get("processes_per_node") or 1)
        has_mpi = bool(mi.get("mpi"))
        world = n * ppn if has_mpi else n
        if mi.get("coordination_command"):
            items.append(("coordination_cmd", mi["coordination_command"]))
        if mi.get("pre_execution_command"):
            items.append(("user_prologue", mi["pre_execution_command"]))
        items += [("num_instances", str(n)), ("ranks_per_instance", str(ppn if has_mpi else 1)), ("world", str(world)),
                  ("master_only", "0" if has_mpi else "1")]
        gpu_list = [n_["gpu_index"] for n_ in nodes if n_.get("gpu_index") is not None]
        for r in range(world):
            items.append(("gpu", str(gpu_list[r % len(gpu_list)]) if gpu_list else "-1"))
        preload_shim = has_mpi and t.get("preload_shim", True)
        env["SHIPYARD_MPI_RUNTIME"] = (mi.get("mpi") or {}).get("runtime", "") if has_mpi else ""
    else:
        items += [("num_instances", "1"), ("ranks_per_instance", "1"), ("world", "1"), ("master_only", "0")]
        g = nodes[0].get("gpu_index") if nodes else None
        items.append(("gpu", "-1" if g is None else str(g)))
    if preload_shim:
        shim = os.path.join(native_dir(), "libshipyard_preload.so")
        if os.path.exists(shim):
            items.append(("preload", shim))
    if t.get("max_wall_time_s"):
        items.append(("wall_time_s", str(int(float(t["max_wall_time_s"])))))
    items.append(("user_cmd", command))
    items += sandbox_items(b, pool, t, tdir, env)
    for k, v in env.items():
        items.append(("env", f"{k}={v}"))
    spec = os.path.join(tdir, "task.spec")
    _write_spec(spec, items)
    with open(os.path.join(tdir, "task.json"), "w") as f:
        json.dump({"task": {k: v for k, v in t.items() if not k.startswith("_")}, "nodes": [n["id"] for n in nodes]}, f,
                  indent=1, default=str)
    return spec, tdir


def build_aux_spec(b, pool: dict, job: dict, kind: str, command: str, node: dict) -> tuple[str, str]:
    """Spec for a job preparation / release command on one node."""
    pid, jid = pool["id"], job["id"]
    tdir = os.path.join(b.job_dir(pid, jid), f"{kind}-{node['id']}")
    wd = os.path.join(tdir, "wd")
    os.makedirs(wd, exist_ok=True)
    env = base_env(b, pool, job, kind, tdir, [node])
    for k, v in (job.get("env") or {}).items():
        env[k] = _expand(str(v), env)
    items = [("workdir", wd), ("taskdir", tdir), ("stdout", os.path.join(tdir, "stdout.txt")),
             ("stderr", os.path.join(tdir, "stderr.txt")), ("result_file", os.path.join(tdir, "result.json")),
             ("world", "1"), ("user_cmd", command)]
    if kind == "jobpreparation":
        pro, _ = data_commands(b, {"input_data": job.get("input_data")}, env)
        if pro:
            items.append(("system_prologue", "set -e; " + "; ".join(pro)))
    for k, v in env.items():
        items.append(("env", f"{k}={v}"))
    spec = os.path.join(tdir, "task.spec")
    _write_spec(spec, items)
    return spec, tdir
