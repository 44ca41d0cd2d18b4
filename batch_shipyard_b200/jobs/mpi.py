"""``multi_instance.mpi`` handling: classic launcher line (dry-run parity) + native launch plan.

The reference turns the ``mpi`` block into an ``mpirun``/``mpiexec`` command
line in the dialect of the chosen runtime (/root/reference/convoy/batch.py:
4362-4486).  We keep that synthesis (table-driven, so a dry run shows the line
a user of the reference expects), but the local backend does not exec it: the
native task runner spawns ``num_instances x processes_per_node`` ranks itself,
one per GPU, and preloads the collectives shim, so ``mpirun`` need not exist.
"""
from __future__ import annotations

import subprocess
from dataclasses import dataclass, field
from typing import Optional, Union

RUNTIMES = ("intelmpi", "intelmpi-ofa", "mpich", "mvapich", "openmpi")

# flag dialects: (host list flag, per-node flag template, leading flags)
_DIALECT = {
    "intelmpi": ("-hosts", "-perhost {ppn}", []),
    "intelmpi-ofa": ("-hosts", "-perhost {ppn}", []),
    "mpich": ("-hosts", "-ppn {ppn}", []),
    "mvapich": ("-hosts", "-ppn {ppn}", []),
    "openmpi": ("-host", "--map-by ppr:{ppn}:node", ["--oversubscribe"]),
}
_PKEY_FILE = "$AZ_BATCH_NODE_STARTUP_DIR/wd/UCX_IB_PKEY"


@dataclass
class MpiSettings:
    runtime: str
    executable_path: str = "mpirun"
    options: list = field(default_factory=list)
    processes_per_node: Union[int, str, None] = 1      # None: not given -> the launcher line carries no -np / per-node flags


def mpi_settings(spec: Optional[dict]) -> Optional[MpiSettings]:
    if not spec:
        return None
    rt = str(spec["runtime"]).lower()
    if rt not in RUNTIMES:
        raise ValueError(f"mpi.runtime '{rt}' is not one of {list(RUNTIMES)}")
    ppn = spec.get("processes_per_node")
    if ppn is not None and not isinstance(ppn, (int, str)) or isinstance(ppn, bool):
        raise ValueError(f"mpi.processes_per_node must be an integer or a command string, not {ppn!r}")
    if isinstance(ppn, str) and ppn.strip().isdigit():
        ppn = int(ppn)
    if isinstance(ppn, int) and ppn < 1:
        raise ValueError("mpi.processes_per_node must be >= 1")
    # NOTE: the executable path keeps its case (the reference lower-cases it, which breaks
    # case-sensitive paths — SURVEY.md Appendix C, Q11)
    return MpiSettings(rt, spec.get("executable_path") or "mpirun", list(spec.get("options") or []), ppn)


def construct_mpi_command(mpi: MpiSettings, num_instances: int, command: str, infiniband: bool = False,
                          rdma_class: str = "none", singularity: Optional[dict] = None,
                          is_docker: bool = True) -> tuple[str, dict]:
    """Returns (launcher command line, extra environment for the fabric).

    ``rdma_class``: 'none' | 'sriov' | 'networkdirect' | 'nvlink' (local NVSwitch box).
    ``processes_per_node`` may be a shell command (e.g. ``nvidia-smi -L | wc -l``); it is
    then evaluated on the node by the shell through ``$(...)``.
    """
    host_flag, per_node, lead = _DIALECT[mpi.runtime]
    ppn = mpi.processes_per_node
    opts = list(mpi.options) + list(lead)
    opts.append(f"{host_flag} $AZ_BATCH_HOST_LIST")
    if ppn is None:
        pass                                       # optional in the reference: the MPI runtime decides
    elif isinstance(ppn, int):
        opts.append(f"-np {num_instances * ppn}")
        opts.append(per_node.format(ppn=ppn))
    else:
        opts.append(f"-np $(expr {num_instances} \\* $({ppn}))")
        opts.append(per_node.format(ppn=f"$({ppn})"))
    env: dict = {}
    if mpi.runtime.startswith("intelmpi") and infiniband:
        env["I_MPI_FALLBACK"] = "0"
        env["MANPATH"] = "/usr/share/man:/usr/local/man"
        if rdma_class == "networkdirect":
            env.update({"I_MPI_FABRICS": "shm:dapl", "I_MPI_DAPL_PROVIDER": "ofa-v2-ib0",
                        "I_MPI_DYNAMIC_CONNECTION": "0", "I_MPI_DAPL_TRANSLATION_CACHE": "0"})
        elif rdma_class == "sriov":
            if mpi.runtime == "intelmpi-ofa":
                env["I_MPI_FABRICS"] = "shm:ofa"
            else:
                env.update({"I_MPI_FABRICS": "shm:ofi", "FI_PROVIDER": "mlx"})
    elif mpi.runtime in ("mpich", "mvapich"):
        if infiniband and rdma_class == "sriov":
            opts.append(f"-env $(cat {_PKEY_FILE})")
    elif mpi.runtime == "openmpi":
        if infiniband and rdma_class == "sriov":
            opts += ["--mca pml ucx", "--mca btl ^vader,tcp,openib", "-x UCX_NET_DEVICES=mlx5_0:1",
                     f"-x $(cat {_PKEY_FILE})"]
        elif rdma_class != "nvlink":
            opts.append("--mca btl_tcp_if_include eth0")
    if rdma_class == "nvlink":
        # one NVSwitch box: the fabric is NVLink and collectives resolve to the preload shim
        env["SHIPYARD_COLL_TRANSPORT"] = "auto"
    if singularity:
        inner = "singularity {} {} {} {}".format(singularity.get("cmd", "exec"),
                                                 " ".join(singularity.get("run_options", [])),
                                                 singularity["image"], command).replace("  ", " ")
        return f"{mpi.executable_path} {' '.join(opts)} {inner}", env
    if mpi.runtime == "openmpi" and is_docker:
        opts.append("--allow-run-as-root")
    return f"{mpi.executable_path} {' '.join(opts)} {command}", env


def resolve_processes_per_node(ppn: Union[int, str], gpu_count: int, dry_run: bool = False) -> int:
    """Evaluate ``processes_per_node`` locally.  Known GPU/CPU counting idioms are answered
    from the topology (so they work in dry-run / on a CPU box); anything else is run by a shell."""
    if ppn is None:
        return 1                                   # the native runner starts one rank per instance
    if isinstance(ppn, int):
        return ppn
    s = ppn.strip()
    norm = " ".join(s.split())
    if norm in ("nvidia-smi -L | wc -l", "nvidia-smi --list-gpus | wc -l"):
        return max(1, gpu_count)
    if dry_run:
        if norm == "nproc":
            import os
            return os.cpu_count() or 1
        return 1
    try:
        out = subprocess.run(["/bin/bash", "-c", s], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                             timeout=30).stdout.strip()
        return max(1, int(out.split()[0]))
    except Exception as e:  # noqa: BLE001
        raise ValueError(f"mpi.processes_per_node command {ppn!r} did not produce an integer: {e}") from e


@dataclass
class LaunchPlan:
    """What the native runner executes for a multi-instance task."""
    world_size: int
    ranks_per_instance: int
    num_instances: int
    gpu_of_rank: list            # local GPU index per rank (or -1 on a CPU pool)
    preload_shim: bool
    shim_face: str               # 'mpi' | 'nccl' | 'none'


def make_launch_plan(num_instances: int, ppn: int, gpus: list, use_shim: bool, has_mpi_block: bool) -> LaunchPlan:
    world = num_instances * ppn
    if gpus:
        gmap = [gpus[r % len(gpus)] for r in range(world)]
    else:
        gmap = [-1] * world
    return LaunchPlan(world, ppn, num_instances, gmap, use_shim, "mpi" if has_mpi_block else ("nccl" if use_shim else "none"))
