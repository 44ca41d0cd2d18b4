"""Federation constraints: parsing + the ordered hard-constraint filter.

Behavioural parity with /root/reference/federation/federation.py: constraint model
(:184-258, memory normalised to MB :212-227), the ordered hard-constraint checks
(:1709-1937) and the node-level checks (low-priority / active-task backlog ratio,
:1939-1996).  Expressed as a rule table: each rule returns ``None`` (pass) or a
``(wanted, actual)`` pair naming why the pool is rejected, so `fed jobs list --blocked`
can show the first failed constraint per pool.

Local meaning of a "pool": a subset of the box's GPUs (1/2/4/8) with its own slots.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Optional

from ..config import settings as S
from ..utils import util


@dataclass
class PoolConstraints:
    autoscale_allow: Optional[bool] = None
    autoscale_exclusive: bool = False
    low_priority_nodes_allow: Optional[bool] = None
    low_priority_nodes_exclusive: bool = False
    native: Optional[bool] = None
    windows: bool = False
    location: Optional[str] = None
    custom_image_arm_id: Optional[str] = None
    virtual_network_arm_id: Optional[str] = None
    registries: list = field(default_factory=list)
    max_active_task_backlog_ratio: Optional[float] = None
    max_active_task_backlog_autoscale_exempt: bool = True


@dataclass
class ComputeNodeConstraints:
    vm_size: Optional[str] = None
    cores: Optional[int] = None
    core_variance: Optional[float] = None
    memory: Optional[float] = None          # MB
    memory_variance: Optional[float] = None
    exclusive: bool = False
    gpu: Optional[bool] = None
    infiniband: Optional[bool] = None


@dataclass
class TaskConstraints:
    auto_complete: bool = False
    has_multi_instance: bool = False
    has_task_dependencies: bool = False
    instance_counts_max: int = 1
    instance_counts_total: int = 1
    merge_task_id: Optional[str] = None
    tasks_per_recurrence: Optional[int] = None


@dataclass
class Constraints:
    pool: PoolConstraints
    compute_node: ComputeNodeConstraints
    task: TaskConstraints


def parse_constraints(jobspec: dict, task_records: list[dict]) -> Constraints:
    fc = jobspec.get("federation_constraints") or {}
    p, cn = fc.get("pool") or {}, fc.get("compute_node") or {}
    if (p.get("autoscale") or {}).get("exclusive") and (p.get("autoscale") or {}).get("allow") is False:
        raise ValueError("federation_constraints.pool.autoscale: exclusive needs allow")
    if (p.get("low_priority_nodes") or {}).get("exclusive") and (p.get("low_priority_nodes") or {}).get("allow") is False:
        raise ValueError("federation_constraints.pool.low_priority_nodes: exclusive needs allow")
    reg = p.get("container_registries") or {}
    registries = list(reg.get("public") or []) + (["hub-private"] if reg.get("private_docker_hub") else [])
    pc = PoolConstraints(
        autoscale_allow=(p.get("autoscale") or {}).get("allow"), autoscale_exclusive=bool((p.get("autoscale") or {}).get("exclusive", False)),
        low_priority_nodes_allow=(p.get("low_priority_nodes") or {}).get("allow"),
        low_priority_nodes_exclusive=bool((p.get("low_priority_nodes") or {}).get("exclusive", False)),
        native=p.get("native"), windows=bool(p.get("windows", False)), location=p.get("location"),
        custom_image_arm_id=(p.get("custom_image_arm_id") or "").lower() or None,
        virtual_network_arm_id=(p.get("virtual_network_arm_id") or "").lower() or None, registries=registries,
        max_active_task_backlog_ratio=(p.get("max_active_task_backlog") or {}).get("ratio"),
        max_active_task_backlog_autoscale_exempt=bool((p.get("max_active_task_backlog") or {}).get("autoscale_exempt", True)))
    cores, mem = cn.get("cores") or {}, cn.get("memory") or {}
    cc = ComputeNodeConstraints(
        vm_size=(cn.get("vm_size") or "").lower() or None, cores=cores.get("amount"), core_variance=cores.get("schedulable_variance"),
        memory=util.parse_size_to_mb(mem.get("amount")) if mem.get("amount") is not None else None,
        memory_variance=mem.get("schedulable_variance"), exclusive=bool(cn.get("exclusive", False)), gpu=cn.get("gpu"),
        infiniband=cn.get("infiniband"))
    if cc.cores is not None and cc.cores <= 0:
        raise ValueError("federation_constraints.compute_node.cores.amount must be positive")
    if cc.memory is not None and cc.memory <= 0:
        raise ValueError("federation_constraints.compute_node.memory.amount must be positive")
    inst = [int((t.get("multi_instance") or {}).get("num_instances") or 1) for t in task_records]
    merge = next((t["id"] for t in task_records if t.get("is_merge_task")), None)
    tc = TaskConstraints(auto_complete=bool(jobspec.get("auto_complete", False)),
                         has_multi_instance=any(t.get("multi_instance") for t in task_records),
                         has_task_dependencies=any(t.get("depends_on") or t.get("depends_on_range") for t in task_records),
                         instance_counts_max=max(inst, default=1), instance_counts_total=sum(inst), merge_task_id=merge)
    return Constraints(pc, cc, tc)


@dataclass
class PoolView:
    """Everything the scheduler needs to know about one candidate pool."""
    id: str
    valid: bool
    location: str
    vm_size: str
    native: bool
    windows: bool
    autoscale_enabled: bool
    target_low_priority: int
    max_tasks_per_node: int
    inter_node_communication: bool
    cores_per_node: int
    memory_mb_per_node: float
    registries: list
    custom_image_arm_id: Optional[str] = None
    virtual_network_arm_id: Optional[str] = None
    idle_dedicated: int = 0
    idle_low_priority: int = 0
    schedulable_dedicated: int = 0
    schedulable_low_priority: int = 0
    active_tasks: int = 0
    blackout_until: float = 0.0


def _node_props(ngpus_total: int) -> tuple[int, float]:
    cores = max(1, (os.cpu_count() or 1) // max(1, ngpus_total))
    try:
        mem = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / (1 << 20) / max(1, ngpus_total)
    except (ValueError, OSError):
        mem = 0.0
    return cores, mem


def pool_view(b, pool_id: str, blackout_until: float = 0.0) -> PoolView:
    p = b.get_pool(pool_id)
    nodes = b.list_nodes(pool_id)
    cores, mem = _node_props(max(1, len(nodes)))
    sched = [n for n in nodes if n["state"] in ("idle", "running")]
    idle = [n for n in nodes if n["state"] == "idle"]
    active = sum(b.count_tasks(j["id"])["active"] for j in b.list_jobs(pool_id) if j["state"] == "active")
    return PoolView(
        id=pool_id, valid=p["state"] == "active" and p["allocation_state"] in ("steady", "resizing"), location="local",
        vm_size=str(p["vm_size"]).lower(), native=bool(p.get("native")), windows=False,
        autoscale_enabled=bool((p.get("autoscale") or {}).get("enabled")), target_low_priority=int(p.get("target_low_priority") or 0),
        max_tasks_per_node=int(p["max_tasks_per_node"]), inter_node_communication=bool(p.get("inter_node_communication_enabled")),
        cores_per_node=cores, memory_mb_per_node=mem, registries=list((p.get("metadata") or {}).get("registries") or []),
        idle_dedicated=sum(1 for n in idle if n["dedicated"]), idle_low_priority=sum(1 for n in idle if not n["dedicated"]),
        schedulable_dedicated=sum(1 for n in sched if n["dedicated"]), schedulable_low_priority=sum(1 for n in sched if not n["dedicated"]),
        active_tasks=active, blackout_until=blackout_until)


Rule = Callable[[PoolView, Constraints], Optional[tuple]]


def _variance(want, var, have) -> Optional[tuple]:
    if want is None:
        return None
    if want > have:
        return (want, have)
    if var == 0 and want != have:
        return (f"=={want}", have)
    if var is not None and var > 0 and have > want * (1 + var):
        return (f"<={want * (1 + var):g}", have)
    return None


HARD_RULES: list[tuple[str, Rule]] = [
    ("valid", lambda p, c: None if p.valid else ("valid pool", "invalid")),
    ("location", lambda p, c: None if not c.pool.location or c.pool.location == p.location else (c.pool.location, p.location)),
    ("virtual_network_arm_id", lambda p, c: None if not c.pool.virtual_network_arm_id or c.pool.virtual_network_arm_id == (p.virtual_network_arm_id or "") else (c.pool.virtual_network_arm_id, p.virtual_network_arm_id)),
    ("custom_image_arm_id", lambda p, c: None if not c.pool.custom_image_arm_id or c.pool.custom_image_arm_id == (p.custom_image_arm_id or "") else (c.pool.custom_image_arm_id, p.custom_image_arm_id)),
    ("windows", lambda p, c: None if not c.pool.windows or p.windows else (True, p.windows)),
    ("native", lambda p, c: None if c.pool.native is None or c.pool.native == p.native else (c.pool.native, p.native)),
    ("autoscale_allow", lambda p, c: None if c.pool.autoscale_allow is not False or not p.autoscale_enabled else (False, True)),
    ("autoscale_exclusive", lambda p, c: None if not c.pool.autoscale_exclusive or p.autoscale_enabled else (True, False)),
    ("low_priority_nodes_allow", lambda p, c: None if c.pool.low_priority_nodes_allow is not False or p.target_low_priority == 0 else (False, p.target_low_priority)),
    ("low_priority_nodes_exclusive", lambda p, c: None if not c.pool.low_priority_nodes_exclusive or p.target_low_priority > 0 or p.autoscale_enabled else (True, 0)),
    ("exclusive", lambda p, c: None if not c.compute_node.exclusive or p.max_tasks_per_node == 1 else (True, p.max_tasks_per_node)),
    ("vm_size", lambda p, c: None if not c.compute_node.vm_size or c.compute_node.vm_size == p.vm_size else (c.compute_node.vm_size, p.vm_size)),
    ("gpu", lambda p, c: None if c.compute_node.gpu is None or c.compute_node.gpu == S.is_gpu_pool(p.vm_size) else (c.compute_node.gpu, S.is_gpu_pool(p.vm_size))),
    ("infiniband", lambda p, c: None if c.compute_node.infiniband is None or c.compute_node.infiniband == S.is_rdma_pool(p.vm_size) else (c.compute_node.infiniband, S.is_rdma_pool(p.vm_size))),
    ("cores", lambda p, c: _variance(c.compute_node.cores, c.compute_node.core_variance, p.cores_per_node)),
    ("memory", lambda p, c: _variance(c.compute_node.memory, c.compute_node.memory_variance, p.memory_mb_per_node) if p.memory_mb_per_node else None),
    ("has_multi_instance", lambda p, c: None if not c.task.has_multi_instance or p.inter_node_communication else (True, False)),
    ("registries", lambda p, c: None if not c.pool.registries or all(r in p.registries for r in c.pool.registries) else (c.pool.registries, p.registries)),
]


def first_failed_hard_constraint(pool: PoolView, c: Constraints) -> Optional[tuple]:
    """(rule name, wanted, actual) of the first violated hard constraint, or None when the pool passes."""
    for name, rule in HARD_RULES:
        r = rule(pool, c)
        if r is not None:
            return (name, r[0], r[1])
    return None


def fails_node_constraints(pool: PoolView, c: Constraints) -> Optional[tuple]:
    """Low-priority and backlog checks that depend on current node counts."""
    if c.pool.low_priority_nodes_allow is False and pool.schedulable_low_priority > 0:
        return ("low_priority_nodes_allow", False, pool.schedulable_low_priority)
    if c.pool.low_priority_nodes_exclusive and pool.schedulable_dedicated > 0 and not pool.autoscale_enabled:
        return ("low_priority_nodes_exclusive", True, pool.schedulable_dedicated)
    ratio = c.pool.max_active_task_backlog_ratio
    if ratio is not None:
        if pool.autoscale_enabled and c.pool.max_active_task_backlog_autoscale_exempt:
            return None
        slots = (pool.schedulable_dedicated + pool.schedulable_low_priority) * pool.max_tasks_per_node
        backlog = pool.active_tasks / slots if slots > 0 else (float("inf") if pool.active_tasks else 0.0)
        if backlog > ratio:
            return ("max_active_task_backlog", ratio, backlog)
    return None
