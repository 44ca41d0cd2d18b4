"""Pool provisioning: the pool-add orchestration + the node ("box") preparation.

Client half — /root/reference/convoy/fleet.py: ``_adjust_settings_for_pool_creation``
cross-field rules (:2637-2901), ``_construct_pool_object`` (nodeprep flag string
:1399-1439, start-task env :1613-1764), ``_add_pool`` (:1821-1940), the wait-for-ready
state machine with reboot / unusable recovery (convoy/batch.py:625-830).
Node half — /root/reference/scripts/shipyard_nodeprep.sh: 26 getopts flags (:127-239),
driver/GPU checks (:626-873), idempotency markers ``.batch_shipyard_node_prep_finished`` /
``.batch_shipyard_cascade_failed`` (:34-36,1935-1973), timing events (:1708-1715),
cascade hand-off (:1669-1750).

Here a node is a GPU of the box: "prep" verifies the GPU through the native probe
(visible, P2P to every peer, NVLS capability), creates the node/shared/startup
directories, runs ``additional_node_prep`` commands, writes the same idempotency
markers, starts exporters, and then runs the cascade pre-load.
"""
from __future__ import annotations

import os
import subprocess
import time
from typing import Optional

from ..backend.local import MAX_REBOOT_RETRIES, LocalBackend
from ..config import settings as S
from ..utils import util
from . import cascade as C
from . import topology

NODEPREP_FINISHED = ".batch_shipyard_node_prep_finished"
CASCADE_FAILED = ".batch_shipyard_cascade_failed"
logger = util.get_logger()


class PoolCreationError(RuntimeError):
    pass


# ---------------------------------------------------------------------------
# cross-field sanity rules applied at `pool add`
# ---------------------------------------------------------------------------
def adjust_settings_for_pool_creation(config: dict) -> list[str]:
    """Validate/adjust pool settings; returns warnings, raises ValueError on hard conflicts."""
    ps = S.pool_settings(config)
    gs = S.global_settings(config)
    warns: list[str] = []
    if len(ps.id) > 64 or not ps.id.replace("-", "").replace("_", "").isalnum():
        raise ValueError("pool id must be 1-64 characters of letters, digits, '-' and '_'")
    if ps.inter_node_communication_enabled and ps.vm_dedicated > 0 and ps.vm_low_priority > 0:
        raise ValueError("inter_node_communication_enabled cannot be combined with both dedicated and low_priority nodes")
    gluster = [v for v in gs.shared_data_volumes.values() if v.volume_driver == "glusterfs_on_compute"]
    if gluster:
        if not ps.inter_node_communication_enabled:
            raise ValueError("glusterfs_on_compute needs inter_node_communication_enabled")
        if ps.vm_low_priority > 0 or ps.vm_dedicated < 2:
            raise ValueError("glusterfs_on_compute needs at least 2 dedicated nodes and no low_priority nodes")
        if ps.autoscale is not None:
            raise ValueError("glusterfs_on_compute cannot be used with autoscale")
        if ps.max_tasks_per_node != 1:
            raise ValueError("glusterfs_on_compute needs max_tasks_per_node: 1")
        for v in gluster:
            if (v.raw.get("volume_type") or "replica") != "replica":
                raise ValueError("glusterfs_on_compute supports only volume_type: replica")
    if ps.per_job_auto_scratch:
        if not ps.inter_node_communication_enabled:
            raise ValueError("per_job_auto_scratch needs inter_node_communication_enabled")
        if ps.is_windows:
            raise ValueError("per_job_auto_scratch is Linux only")
    if ps.is_windows and (ps.node_exporter_enabled or ps.cadvisor_enabled):
        raise ValueError("prometheus exporters are Linux only")
    if (gs.singularity_images_unsigned or gs.singularity_images_signed) and (ps.native or ps.is_windows):
        raise ValueError("singularity images cannot be used with native or Windows pools")
    if "kata_containers" in ps.container_runtimes_install:
        warns.append("kata_containers needs nested virtualisation; ignored on a local GPU box")
    if ps.custom_image and ps.attempt_recovery_on_unusable:
        S.set_attempt_recovery_on_unusable(config, False)
        warns.append("attempt_recovery_on_unusable forced off for custom images")
    if ps.transfer_files_on_pool_creation and ps.block_until_all_global_resources_loaded:
        S.set_block_until_all_global_resources_loaded(config, False)
        warns.append("block_until_all_global_resources_loaded forced off because transfer_files_on_pool_creation is set")
    if ps.is_windows:
        warns.append("Windows node scripts are out of scope on a Linux GPU box; treating as Linux")
    if S.is_gpu_pool(ps.vm_size) and not S.is_gpu_compute_pool(ps.vm_size):
        warns.append("visualisation GPU sizes have no NVLink collectives path")
    return warns


def nodeprep_flags(config: dict) -> str:
    """The flag string a node-prep invocation receives (shown by dry-run / `--show-config`)."""
    ps, gs = S.pool_settings(config), S.global_settings(config)
    f = []
    if ps.block_until_all_global_resources_loaded:
        f.append("-b")
    if S.is_gpu_pool(ps.vm_size):
        f.append("-g " + (ps.gpu_driver_source or "box-driver"))
    if ps.gpu_ignore_warnings:
        f.append("-i")
    if ps.inter_node_communication_enabled:
        f.append("-c")
    if ps.per_job_auto_scratch:
        f.append("-j")
    if gs.delay_docker_image_preload:
        f.append("-d")
    if gs.fallback_registry:
        f.append(f"-l {gs.fallback_registry}")
    if gs.store_timing_metrics:
        f.append("-p")
    if ps.node_exporter_enabled:
        f.append(f"-q ne:{ps.node_exporter_port}")
    if ps.cadvisor_enabled:
        f.append(f"-q ca:{ps.cadvisor_port}")
    f.append(f"-r {ps.container_runtimes_default}")
    f.append(f"-n {'native' if ps.native else 'shipyard'}")
    f.append(f"-s {gs.storage_entity_prefix}")
    f.append(f"-o {gs.concurrent_source_downloads}")
    return " ".join(f)


# ---------------------------------------------------------------------------
# node preparation
# ---------------------------------------------------------------------------
def _marker(b: LocalBackend, pool_id: str, node_id: str, name: str) -> str:
    d = os.path.join(b.node_startup_dir(pool_id), node_id)
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, name)


def prep_node(b: LocalBackend, pool_id: str, node: dict, ps: Optional[S.PoolSettings] = None,
              gpus: Optional[list] = None, timing: bool = True) -> tuple[bool, str]:
    """Run the start-task equivalent for one node.  Returns (ok, message)."""
    nid = node["id"]
    fin = _marker(b, pool_id, nid, NODEPREP_FINISHED)
    if os.path.exists(fin):
        return True, "node prep already finished (reboot fast-path)"
    if timing:
        b.store.record_event("nodeprep", "start", pool=pool_id, node=nid)
    b.set_node_state(pool_id, nid, "waiting_for_start_task")
    gi = node.get("gpu_index")
    if gi is not None:
        inv = {g["index"]: g for g in (gpus if gpus is not None else topology.probe_gpus())}
        g = inv.get(gi)
        if g is None:
            return False, f"unusable: GPU {gi} is not visible to the driver"
        bad = [p["peer"] for p in g.get("p2p", []) if not p.get("access")]
        ignore = bool(ps and ps.gpu_ignore_warnings)
        if bad and not ignore:
            return False, f"GPU {gi} has no peer access to GPU(s) {bad} (NVSwitch fabric down?); set gpu.ignore_warnings to continue"
    stdout = open(_marker(b, pool_id, nid, "stdout.txt"), "a")
    env = dict(os.environ)
    env.update({"AZ_BATCH_POOL_ID": pool_id, "AZ_BATCH_NODE_ID": nid, "AZ_BATCH_NODE_ROOT_DIR": b.pool_root(pool_id),
                "AZ_BATCH_NODE_SHARED_DIR": b.node_shared_dir(pool_id), "AZ_BATCH_NODE_STARTUP_DIR": b.node_startup_dir(pool_id),
                "SHIPYARD_GPU": "" if gi is None else str(gi)})
    if ps is not None:
        env.update({str(k): str(v) for k, v in ps.additional_node_prep_env.items()})
        for cmd in list(ps.additional_node_prep_pre) + list(ps.additional_node_prep_post):
            rc = subprocess.call(["/bin/bash", "-c", cmd], env=env, stdout=stdout, stderr=subprocess.STDOUT,
                                 cwd=b.node_startup_dir(pool_id))
            if rc != 0:
                stdout.close()
                return False, f"additional_node_prep command failed ({rc}): {cmd}"
    stdout.close()
    with open(fin, "w") as f:
        f.write(util.datetime_utcnow(as_string=True) + "\n")
    if timing:
        b.store.record_event("nodeprep", "end", pool=pool_id, node=nid)
    return True, "ok"


def bring_up_nodes(b: LocalBackend, pool_id: str, ps: Optional[S.PoolSettings] = None, gpus: Optional[list] = None,
                   fault_hook=None) -> dict:
    """Start-task + recovery state machine over all nodes not yet ready.

    start_task_failed -> reboot (re-run prep) up to 5 times when ``reboot_on_start_task_failed``;
    unusable -> delete the node and re-resize when ``attempt_recovery_on_unusable``.
    ``fault_hook(node, attempt) -> Optional[str]`` lets tests inject failures.
    """
    pool = b.get_pool(pool_id)
    summary = {"ready": 0, "start_task_failed": 0, "unusable": 0, "rebooted": 0, "recovered": 0}
    reboot_ok = bool(ps.reboot_on_start_task_failed) if ps else bool(pool.get("reboot_on_start_task_failed"))
    recover_ok = bool(ps.attempt_recovery_on_unusable) if ps else bool(pool.get("attempt_recovery_on_unusable"))
    for node in b.list_nodes(pool_id):
        if node["state"] in ("idle", "running", "preempted"):
            summary["ready"] += 1
            continue
        if node["state"] not in ("creating", "starting", "rebooting", "start_task_failed", "waiting_for_start_task", "unusable"):
            continue
        attempt = 0
        while True:
            b.set_node_state(pool_id, node["id"], "starting")
            t_prep = time.time()
            injected = fault_hook(node, attempt) if fault_hook else None
            ok, msg = (False, injected) if injected else prep_node(b, pool_id, node, ps, gpus)
            if ok:
                b.set_node_state(pool_id, node["id"], "idle", last_boot_time=time.time(), start_task_seconds=time.time() - t_prep,
                                 start_task={"exit_code": 0, "message": msg, "attempts": attempt + 1})
                summary["ready"] += 1
                break
            if msg.startswith("unusable"):
                b.set_node_state(pool_id, node["id"], "unusable", errors=[msg])
                summary["unusable"] += 1
                if recover_ok and attempt < 1:
                    # delete the node and allocate a replacement
                    tgt = b.get_pool(pool_id)
                    b.store.delete("node", pool_id, node["id"])
                    b.resize_pool(pool_id, tgt["target_dedicated"], tgt["target_low_priority"])
                    summary["recovered"] += 1
                break
            b.set_node_state(pool_id, node["id"], "start_task_failed", errors=[msg],
                             start_task={"exit_code": 1, "message": msg, "attempts": attempt + 1})
            if reboot_ok and attempt < MAX_REBOOT_RETRIES:
                attempt += 1
                summary["rebooted"] += 1
                b.set_node_state(pool_id, node["id"], "rebooting", reboots=attempt)
                continue
            summary["start_task_failed"] += 1
            break
    nodes = b.list_nodes(pool_id)
    steady = all(n["state"] not in ("creating", "starting", "rebooting", "waiting_for_start_task") for n in nodes)
    if steady:
        b.set_pool_allocation_state(pool_id, "steady")
    return summary


def run_cascade(b: LocalBackend, config: dict, pool_id: str, block: bool = True, device: Optional[int] = None) -> bool:
    gs = S.global_settings(config)
    C.Cascade.populate(b.store, pool_id, S.global_resources_images(config))
    cas = C.Cascade(b.store, pool_id, concurrency=gs.concurrent_source_downloads, device=device,
                    fallback_registry=gs.fallback_registry)
    ok = cas.run(block=block)
    if not ok:
        with open(os.path.join(b.node_startup_dir(pool_id), CASCADE_FAILED), "w") as f:
            f.write("\n".join(f"{k}: {v}" for k, v in cas.errors.items()))
    return ok


# ---------------------------------------------------------------------------
# `pool add`
# ---------------------------------------------------------------------------
def create_pool(b: LocalBackend, config: dict, recreate: bool = False, no_wait: bool = False,
                fault_hook=None) -> dict:
    warns = adjust_settings_for_pool_creation(config)
    ps = S.pool_settings(config)
    for w in warns:
        logger.warning(w)
    if b.pool_exists(ps.id):
        if not recreate:
            raise PoolCreationError(f"pool {ps.id} already exists (use --recreate)")
        b.delete_pool(ps.id)
    gpus_cfg = S.credentials_local_gpus(config)
    inv = topology.probe_gpus()
    gpu_ids = [g["index"] for g in inv]
    if gpus_cfg is not None:
        gpu_ids = [g for g in gpu_ids if g in set(gpus_cfg)]
    if S.is_gpu_pool(ps.vm_size) and not gpu_ids and not ps.gpu_ignore_warnings:
        raise PoolCreationError(f"pool vm_size {ps.vm_size} is a GPU size but no GPU is visible "
                                "(set gpu.ignore_warnings: true to create a CPU-slot pool)")
    want = ps.vm_dedicated + ps.vm_low_priority
    use = gpu_ids[:want] if (gpu_ids and want > 0) else gpu_ids
    formula = None
    if ps.autoscale is not None:
        from . import autoscale as AS
        formula = AS.get_formula(ps)
    # record storage links so task-side movers can resolve them without the config files
    for link, v in ((config.get("credentials") or {}).get("storage") or {}).items():
        if isinstance(v, dict) and link != "aad":
            b.store.insert("storagelink", link, "", {"local_path": v.get("local_path")}, replace=True)
    meta = {"nodeprep_flags": nodeprep_flags(config), "topology": topology.describe()}
    pool = b.create_pool(ps, gpus=use if use else None, metadata=meta)
    if formula:
        b.store.mutate("pool", ps.id, "", lambda p: p["autoscale"].update({"formula": formula, "enabled": True}))
    if no_wait:
        return b.get_pool(ps.id)
    summary = bring_up_nodes(b, ps.id, ps, inv if inv else None, fault_hook=fault_hook)
    nodes = b.list_nodes(ps.id)
    if summary["ready"] == 0 and nodes:
        raise PoolCreationError(f"no node of pool {ps.id} became ready: "
                                + "; ".join(f"{n['id']}: {n['state']} {n.get('errors')}" for n in nodes))
    gs = S.global_settings(config)
    dev = use[0] if use else None
    ok = run_cascade(b, config, ps.id, block=ps.block_until_all_global_resources_loaded, device=dev)
    if not ok and ps.block_until_all_global_resources_loaded:
        raise PoolCreationError(f"global resources failed to load on pool {ps.id}; see {CASCADE_FAILED} in the startup dir")
    out = b.get_pool(ps.id)
    out["_summary"] = summary
    return out
