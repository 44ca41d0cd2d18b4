"""Task construction: jobs.yaml task spec -> concrete task record for the backend.

Covers what the reference does in ``settings.task_settings``
(/root/reference/convoy/settings.py:3727-4552) and ``batch._construct_task``
(/root/reference/convoy/batch.py:4489-4946): task-id allocation, container
run-option synthesis (docker / singularity strings, kept for dry-run parity and
used verbatim when a container runtime exists on the box), GPU / "infiniband"
(= NVLink here) flags, env + env-file exclusion, data/shared volume binds,
user identity, working dir, multi-instance coordination / application phases,
MPI launcher synthesis, dependencies, exit conditions, retries / wall time,
input/output data movement specs.

There is ONE real execution mode on the local box (the native runner owns the
ranks), but both ``native`` and non-native YAML spellings are accepted and the
record says which command shape was synthesised.
"""
from __future__ import annotations

import re
import shlex
from dataclasses import asdict, dataclass, field
from typing import Optional

from ..config import settings as S
from ..utils import util
from . import mpi as M

# environment names never forwarded into the container env-file
ENV_EXCLUDE = ("_", "HOME", "HOSTNAME", "PATH", "PWD", "SHLVL", "USER")
MAX_TASK_ID_LEN = 64
_TASK_ID_RE = re.compile(r"^[a-zA-Z0-9_-]{1,64}$")


@dataclass
class TaskRecord:
    id: str
    job_id: str
    command: Optional[str]
    image: Optional[str]
    runtime: str                       # docker | singularity | process
    native_shape: bool                 # command shape synthesised for a `native` pool
    run_options: list = field(default_factory=list)
    container_command: Optional[str] = None     # full docker/singularity line (dry-run / container runtimes)
    env: dict = field(default_factory=dict)
    env_exclude: list = field(default_factory=lambda: list(ENV_EXCLUDE))
    depends_on: list = field(default_factory=list)
    depends_on_range: Optional[list] = None
    max_task_retries: int = 0
    max_wall_time_s: Optional[float] = None
    retention_time_s: Optional[float] = None
    exit_job_action: str = "none"
    exit_dependency_action: str = "block"
    multi_instance: Optional[dict] = None
    mpi_command: Optional[str] = None
    gpus: Optional[str] = None
    infiniband: bool = False
    working_dir: str = "batch"
    user_identity: Optional[dict] = None
    resource_files: list = field(default_factory=list)
    input_data: Optional[dict] = None
    output_data: Optional[dict] = None
    system_prologue: list = field(default_factory=list)
    system_epilogue: list = field(default_factory=list)
    name: Optional[str] = None
    labels: list = field(default_factory=list)
    is_merge_task: bool = False
    # container semantics the native runner enforces with a process sandbox (mount namespace): see runspec.sandbox_items
    sandbox: Optional[dict] = None

    def to_dict(self) -> dict:
        return asdict(self)


# ---------------------------------------------------------------------------
# task ids
# ---------------------------------------------------------------------------
class TaskIdAllocator:
    """Generic task ids ``<prefix>NNNNN`` continuing after the highest id already in the job.  The highest number per prefix is
    found once (one pass over the existing ids) and then counted up, so a 10 000-task sweep does not rescan the job for every
    task (the reference lists the job's tasks once per submission for the same reason, convoy/batch.py:4165-4240)."""

    def __init__(self, existing: set, reserved: Optional[set] = None):
        self._ids = set(existing) | set(reserved or ())
        self._next: dict = {}

    def _start(self, pfx: str) -> int:
        pat = re.compile("^" + re.escape(pfx) + r"(\d+)$")
        mx = -1
        for tid in self._ids:
            m = pat.match(tid)
            if m:
                mx = max(mx, int(m.group(1)))
        return mx + 1

    def reserve(self, tid: str) -> None:
        self._ids.add(tid)
        for pfx in self._next:                       # an explicit id of the same shape moves the counter past it
            if tid.startswith(pfx) and tid[len(pfx):].isdigit():
                self._next[pfx] = max(self._next[pfx], int(tid[len(pfx):]) + 1)

    def next(self, prefix: str, zfill: int, is_merge: bool = False) -> str:
        pfx = ("merge-" if is_merge else "") + prefix
        if pfx not in self._next:
            self._next[pfx] = self._start(pfx)
        tid = f"{pfx}{str(self._next[pfx]).zfill(zfill)}"
        if len(tid) > MAX_TASK_ID_LEN:
            raise ValueError(f"generated task id '{tid}' exceeds {MAX_TASK_ID_LEN} characters")
        self._next[pfx] += 1
        self._ids.add(tid)
        return tid


def next_generic_task_id(existing: set, prefix: str, zfill: int, reserved: Optional[set] = None,
                         is_merge: bool = False) -> str:
    """One-shot form of :class:`TaskIdAllocator` (``merge-<prefix>NNNNN`` for merge tasks)."""
    return TaskIdAllocator(existing, reserved).next(prefix, zfill, is_merge)


def validate_task_id(tid: str) -> None:
    if not _TASK_ID_RE.match(tid):
        raise ValueError(f"task id '{tid}' must be 1-64 characters of [a-zA-Z0-9_-]")


# ---------------------------------------------------------------------------
# run-option synthesis
# ---------------------------------------------------------------------------
def _gpu_options(task: dict, jobspec: dict, pool: S.PoolSettings, singularity: bool) -> tuple[Optional[str], list]:
    gpus = task.get("gpus", jobspec.get("gpus"))
    if gpus is None:
        gpus = "all" if S.is_gpu_pool(pool.vm_size) else "disable"
    gpus = str(gpus)
    if gpus == "disable":
        return gpus, []
    if not S.is_gpu_pool(pool.vm_size) and not pool.gpu_ignore_warnings:
        raise ValueError(f"task requests gpus={gpus} but pool vm_size {pool.vm_size} has no GPUs "
                         "(set pool gpu.ignore_warnings to override)")
    if singularity:
        if gpus != "all":
            raise ValueError("singularity tasks support only gpus: all (--nv)")
        return gpus, ["--nv"]
    return gpus, [f"--gpus={gpus}"]


def _infiniband(task: dict, jobspec: dict, pool: S.PoolSettings) -> bool:
    v = task.get("infiniband")
    if v is None:
        v = jobspec.get("infiniband")     # job-level value is honoured (the reference drops it, Q4)
    if v is None:
        return S.is_rdma_pool(pool.vm_size) and pool.inter_node_communication_enabled and not pool.is_windows
    return bool(v)


def sandbox_binds(config: dict, task: dict, jobspec: dict) -> list:
    """[source, destination, options] for every data / shared data volume of the task: what the runner bind-mounts inside the
    task's mount namespace (the same volumes `_volume_binds` renders as -v / -B strings).  A data volume without host_path is an
    anonymous volume: source None, materialised under the task's container scratch by the runner spec."""
    gs = S.global_settings(config)
    out = []
    for vname in list(jobspec.get("data_volumes") or []) + list(task.get("data_volumes") or []):
        dv = gs.data_volumes.get(vname)
        if dv is None:
            raise ValueError(f"data volume '{vname}' is not defined in global_resources.volumes.data_volumes")
        out.append([dv.host_path or None, dv.container_path, dv.bind_options or "", vname])
    for vname in list(jobspec.get("shared_data_volumes") or []) + list(task.get("shared_data_volumes") or []):
        sv = gs.shared_data_volumes.get(vname)
        if sv is None:
            raise ValueError(f"shared data volume '{vname}' is not defined in global_resources.volumes.shared_data_volumes")
        out.append([shared_volume_host_path(sv), sv.container_path, sv.bind_options or "", vname])
    return out


def _volume_binds(config: dict, task: dict, jobspec: dict, singularity: bool) -> list:
    gs = S.global_settings(config)
    opts = []
    flag = "-B" if singularity else "-v"
    for vname in list(jobspec.get("data_volumes") or []) + list(task.get("data_volumes") or []):
        dv = gs.data_volumes.get(vname)
        if dv is None:
            raise ValueError(f"data volume '{vname}' is not defined in global_resources.volumes.data_volumes")
        src = dv.host_path
        bind = f"{src}:{dv.container_path}" if src else dv.container_path
        if dv.bind_options:
            bind += f":{dv.bind_options}"
        opts.append(f"{flag} {bind}")
    for vname in list(jobspec.get("shared_data_volumes") or []) + list(task.get("shared_data_volumes") or []):
        sv = gs.shared_data_volumes.get(vname)
        if sv is None:
            raise ValueError(f"shared data volume '{vname}' is not defined in global_resources.volumes.shared_data_volumes")
        host = shared_volume_host_path(sv)
        bind = f"{host}:{sv.container_path}"
        if sv.bind_options:
            bind += f":{sv.bind_options}"
        opts.append(f"{flag} {bind}")
    return opts


def shared_volume_host_path(sv: S.SharedDataVolume) -> str:
    """Where a shared data volume lives on the box (every driver maps to a directory under the shared dir)."""
    if sv.volume_driver == "custom_linux_mount":
        return f"$AZ_BATCH_NODE_ROOT_DIR/mounts/{sv.name}"
    if sv.volume_driver == "glusterfs_on_compute":
        return "$AZ_BATCH_NODE_SHARED_DIR/.gluster/gv0"
    return f"$AZ_BATCH_NODE_ROOT_DIR/mounts/{sv.volume_driver}-{sv.name}"


def _user_identity(jobspec: dict) -> tuple[Optional[dict], list]:
    ui = jobspec.get("user_identity") or {}
    admin, su = ui.get("default_pool_admin"), ui.get("specific_user")
    if admin and su:
        raise ValueError("user_identity: default_pool_admin and specific_user are mutually exclusive")
    if su:
        opts = [f"-u {su['uid']}:{su['gid']}", "-v /etc/passwd:/etc/passwd:ro", "-v /etc/group:/etc/group:ro",
                "-v /etc/sudoers:/etc/sudoers:ro"]
        return {"uid": int(su["uid"]), "gid": int(su["gid"])}, opts
    if admin:
        return {"default_pool_admin": True}, []
    return None, []


def _working_dir(task: dict, jobspec: dict, singularity: bool) -> tuple[str, list]:
    wd = task.get("default_working_dir") or jobspec.get("default_working_dir") or "batch"
    if wd == "batch":
        return wd, (["--pwd $AZ_BATCH_TASK_WORKING_DIR"] if singularity else ["-w $AZ_BATCH_TASK_WORKING_DIR"])
    return wd, []


def _bind_defaults(jobspec: dict, singularity: bool) -> list:
    flag = "-B" if singularity else "-v"
    if bool(jobspec.get("restrict_default_bind_mounts", False)):
        return [f"{flag} $AZ_BATCH_TASK_DIR:$AZ_BATCH_TASK_DIR"]
    return [f"{flag} $AZ_BATCH_NODE_ROOT_DIR:$AZ_BATCH_NODE_ROOT_DIR"]


def _seconds(v) -> Optional[float]:
    td = util.convert_string_to_timedelta(v)
    return None if td is None else td.total_seconds()


# ---------------------------------------------------------------------------
# main entry
# ---------------------------------------------------------------------------
def build_task(config: dict, pool: S.PoolSettings, jobspec: dict, task: dict, task_id: str,
               pool_counts: Optional[dict] = None, is_merge: bool = False, gpu_count: int = 0,
               dry_run: bool = False) -> TaskRecord:
    """Turn one expanded task spec into a TaskRecord."""
    jid = S.job_id(jobspec)
    validate_task_id(task_id)
    docker_image, sing_image = task.get("docker_image"), task.get("singularity_image")
    if docker_image and sing_image:
        raise ValueError(f"task {task_id}: specify docker_image or singularity_image, not both")
    if not docker_image and not sing_image:
        raise ValueError(f"task {task_id}: a docker_image or singularity_image is required")
    singularity = bool(sing_image)
    if singularity and pool.native:
        raise ValueError("singularity images cannot run on native container pools")
    image = sing_image or docker_image
    gs = S.global_settings(config)
    # missing-image policy: a task may only use a preloaded image unless the job allows otherwise
    known = gs.singularity_images_unsigned + gs.singularity_images_signed if singularity else gs.docker_images
    if image not in known and not S.job_allow_run_on_missing_image(jobspec):
        raise ValueError(f"task {task_id}: image '{image}' is not in global_resources "
                         f"({'singularity_images' if singularity else 'docker_images'}); add it there or set "
                         "allow_run_on_missing_image: true on the job")

    mi = task.get("multi_instance")
    run_opts: list = []
    # remove-after-exit: task overrides job, default true
    rm = task.get("remove_container_after_exit", jobspec.get("remove_container_after_exit", True))
    if rm and not singularity and not mi:
        run_opts.append("--rm")
    shm = task.get("shm_size") or jobspec.get("shm_size")
    if shm and not singularity:
        run_opts.append(f"--shm-size={shm}")
    name = task.get("name")
    if not singularity:
        if mi:
            name = util.normalize_docker_image_name_for_job(jid, image)
        run_opts.append(f"--name {name or task_id}")
        for lb in task.get("labels") or []:
            run_opts.append(f"-l {lb}")
        for p in task.get("ports") or []:
            run_opts.append(f"-p {p}")
        if task.get("entrypoint"):
            run_opts.append(f"--entrypoint {task['entrypoint']}")
    run_opts.append("--env-file $AZ_BATCH_TASK_DIR/.shipyard.envlist" if not singularity else "")
    run_opts += _bind_defaults(jobspec, singularity)
    run_opts += _volume_binds(config, task, jobspec, singularity)
    ident, id_opts = _user_identity(jobspec)
    if not singularity:
        run_opts += id_opts
    wd, wd_opts = _working_dir(task, jobspec, singularity)
    run_opts += wd_opts
    gpus, gpu_opts = _gpu_options(task, jobspec, pool, singularity)
    run_opts += gpu_opts
    ib = _infiniband(task, jobspec, pool)
    local_box = S.local_gpu_count_from_vm_size(pool.vm_size) is not None or pool.vm_size.lower().startswith(("b200", "local"))
    if ib and not singularity:
        if local_box:
            # NVSwitch box: "infiniband" means the NVLink fabric — host IPC so ranks can exchange cuMem fds
            run_opts += ["--net=host", "--ipc=host", "--ulimit memlock=-1"]
        else:
            run_opts += ["--net=host", "--ulimit memlock=9223372036854775807", "--device=/dev/infiniband/rdma_cm",
                         "--device=/dev/infiniband/uverbs0"]
    run_opts += list(task.get("additional_singularity_options" if singularity else "additional_docker_run_options") or [])
    run_opts = [o for o in run_opts if o]

    env = dict(S.job_environment_variables(jobspec))
    env.update({str(k): "" if v is None else str(v) for k, v in (task.get("environment_variables") or {}).items()})
    if gpus != "disable":
        env.setdefault("CUDA_CACHE_DISABLE", "0")
        env.setdefault("CUDA_CACHE_MAXSIZE", "1073741824")
        env.setdefault("CUDA_CACHE_PATH", "$AZ_BATCH_NODE_SHARED_DIR/.nv/ComputeCache")

    command = task.get("command")
    rec = TaskRecord(id=task_id, job_id=jid, command=command, image=image,
                     runtime="singularity" if singularity else "docker", native_shape=pool.native,
                     run_options=run_opts, env=env, gpus=gpus, infiniband=ib, working_dir=wd, user_identity=ident,
                     name=name, labels=list(task.get("labels") or []), is_merge_task=is_merge)

    rec.sandbox = {"binds": sandbox_binds(config, task, jobspec),
                   "restrict_default_bind_mounts": bool(jobspec.get("restrict_default_bind_mounts", False)),
                   "remove_after_exit": bool(rm), "shm_size": shm or None, "name": name or task_id,
                   "uid": (ident or {}).get("uid"), "gid": (ident or {}).get("gid")}

    # retries / wall / retention: task overrides job
    rec.max_task_retries = int(task.get("max_task_retries", S.job_max_task_retries(jobspec)) or 0)
    rec.max_wall_time_s = _seconds(task.get("max_wall_time") or jobspec.get("max_wall_time"))
    rec.retention_time_s = _seconds(task.get("retention_time") or jobspec.get("retention_time")) or 7 * 86400.0
    eo = S.exit_options(task, S.exit_options(jobspec))
    rec.exit_job_action, rec.exit_dependency_action = eo.job_action, eo.dependency_action

    # dependencies
    rec.depends_on = [str(x) for x in (task.get("depends_on") or [])]
    dr = task.get("depends_on_range")
    if dr:
        if len(dr) != 2 or not all(isinstance(x, int) for x in dr):
            raise ValueError(f"task {task_id}: depends_on_range needs exactly two integers")
        rec.depends_on_range = [int(dr[0]), int(dr[1])]

    # data movement
    rec.resource_files = list(task.get("resource_files") or [])
    if task.get("input_data"):
        if pool.native:
            raise ValueError(f"task {task_id}: task-level input_data is not supported on native container pools")
        rec.input_data = task["input_data"]
    rec.output_data = task.get("output_data")

    sing_cmd = (task.get("singularity_execution") or {}).get("cmd", "exec")
    # multi-instance
    if mi:
        counts = pool_counts or {"current_dedicated": pool.vm_dedicated, "current_low_priority": pool.vm_low_priority}
        n = S.resolve_num_instances(mi["num_instances"], counts.get("current_dedicated", 0),
                                    counts.get("current_low_priority", 0), pool.vm_dedicated, pool.vm_low_priority)
        coord = mi.get("coordination_command")
        mpi = M.mpi_settings(mi.get("mpi"))
        if singularity:
            coord_line = ":"                      # singularity: no daemonised container
        elif pool.native:
            coord_line = coord or "/usr/sbin/sshd -p 23"
        else:
            # non-native docker: coordination container is daemonised on every instance, never auto-removed
            copts = [o for o in run_opts if o != "--rm"]
            if "--net=host" not in copts:
                copts.append("--net=host")
            coord_line = "docker run -d {} {}{}".format(" ".join(copts), image, f" {coord}" if coord else "")
        ppn_raw = mpi.processes_per_node if mpi else 1
        ppn = M.resolve_processes_per_node(ppn_raw, gpu_count, dry_run=dry_run) if mpi else 1
        rec.multi_instance = {"num_instances": n, "coordination_command": coord, "coordination_line": coord_line,
                              "pre_execution_command": mi.get("pre_execution_command"),
                              "resource_files": list(mi.get("resource_files") or []),
                              "processes_per_node": ppn, "processes_per_node_raw": ppn_raw,
                              "mpi": None if mpi is None else {"runtime": mpi.runtime, "executable_path": mpi.executable_path,
                                                               "options": mpi.options}}
        if mpi:
            rdma = "nvlink" if local_box else ("sriov" if S.is_rdma_pool(pool.vm_size) else "none")
            sing = {"cmd": sing_cmd, "run_options": run_opts, "image": image} if singularity else None
            line, fabric_env = M.construct_mpi_command(mpi, n, command or "", ib, rdma, sing, is_docker=not singularity)
            rec.mpi_command = line
            for k, v in fabric_env.items():
                rec.env.setdefault(k, v)

    # the container command line a runtime would execute (shown by dry-run, used when docker/singularity exist)
    inner = rec.mpi_command if (rec.mpi_command and not singularity) else (command or "")
    if singularity:
        rec.container_command = rec.mpi_command or "singularity {} {} {} {}".format(sing_cmd, " ".join(run_opts), image, command or "").strip()
    elif mi and not pool.native:
        rec.container_command = f"docker exec {' '.join(o for o in ['-e', 'SHIPYARD_TASK=1'] )} {name} {inner}".replace("  ", " ").strip()
    else:
        rec.container_command = f"docker run {' '.join(run_opts)} {image} {inner}".strip()
    return rec


def env_dump_command(exclude=ENV_EXCLUDE, env_file: str = "$AZ_BATCH_TASK_DIR/.shipyard.envlist") -> str:
    """Shell line writing the environment (minus excluded names) to the env-file."""
    if not exclude:
        return f"env > {env_file}"                         # (the reference emits `env | file` here, Q3)
    pat = "|".join(f"^{re.escape(n)}=" for n in exclude)
    return f"env | grep -Ev {shlex.quote(pat)} > {env_file}"
