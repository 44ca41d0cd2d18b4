"""Greedy best-fit pool selection for federated jobs.

Order of preference (/root/reference/federation/federation.py:2030-2228): a pool with enough
**idle** capacity, then enough **available** (schedulable) capacity, then an **autoscale**
pool that can grow, then the least **backlogged** pool.  Multi-instance jobs are matched by
NODES >= max instance count; ordinary jobs by SLOTS >= total tasks (:2164-2170).  Ties go to
the tightest fit (least left-over capacity), which keeps big pools free for big jobs.
"""
from __future__ import annotations

import time
from typing import Optional

from .constraints import Constraints, PoolView, fails_node_constraints, first_failed_hard_constraint


def _need(c: Constraints) -> tuple[bool, int]:
    if c.task.has_multi_instance:
        return True, c.task.instance_counts_max
    return False, c.task.tasks_per_recurrence or c.task.instance_counts_total


def _capacity(p: PoolView, by_nodes: bool, idle: bool, allow_lp: bool, lp_only: bool) -> int:
    d = p.idle_dedicated if idle else p.schedulable_dedicated
    lp = p.idle_low_priority if idle else p.schedulable_low_priority
    nodes = (0 if lp_only else d) + (lp if allow_lp else 0)
    return nodes if by_nodes else nodes * p.max_tasks_per_node


def select_pool(pools: list[PoolView], c: Constraints, now: Optional[float] = None) -> tuple[Optional[str], dict]:
    """Returns (pool id or None, diagnostics {pool id: reason})."""
    now = time.time() if now is None else now
    diag: dict = {}
    cands = []
    for p in pools:
        if p.blackout_until > now:
            diag[p.id] = "blackout after recent scheduling"
            continue
        hf = first_failed_hard_constraint(p, c)
        if hf is not None:
            diag[p.id] = f"hard constraint {hf[0]}: wanted {hf[1]}, pool has {hf[2]}"
            continue
        nf = fails_node_constraints(p, c)
        if nf is not None:
            diag[p.id] = f"node constraint {nf[0]}: wanted {nf[1]}, pool has {nf[2]}"
            continue
        cands.append(p)
    if not cands:
        return None, diag
    by_nodes, need = _need(c)
    allow_lp = c.pool.low_priority_nodes_allow is not False
    lp_only = c.pool.low_priority_nodes_exclusive

    def best(idle: bool):
        fits = [(cap - need, p) for p in cands if (cap := _capacity(p, by_nodes, idle, allow_lp, lp_only)) >= need]
        return min(fits, key=lambda x: (x[0], x[1].id))[1] if fits else None

    for stage, idle in (("idle", True), ("available", False)):
        p = best(idle)
        if p is not None:
            diag[p.id] = f"selected: {stage} capacity"
            return p.id, diag
    auto = [p for p in cands if p.autoscale_enabled and c.pool.autoscale_allow is not False]
    if auto:
        p = min(auto, key=lambda x: (x.active_tasks, x.id))
        diag[p.id] = "selected: autoscale-enabled pool can grow"
        return p.id, diag
    # backlog: only for slot-matched jobs on pools that can eventually run them
    runnable = [p for p in cands if _capacity(p, by_nodes, False, allow_lp, lp_only) > 0 and (not by_nodes or _capacity(p, True, False, allow_lp, lp_only) >= need)]
    if runnable:
        def backlog(p):
            slots = max(1, _capacity(p, False, False, allow_lp, lp_only))
            return (p.active_tasks / slots, p.id)
        p = min(runnable, key=backlog)
        diag[p.id] = "selected: least backlog"
        return p.id, diag
    for p in cands:
        diag.setdefault(p.id, f"insufficient capacity: need {need} {'nodes' if by_nodes else 'slots'}")
    return None, diag
