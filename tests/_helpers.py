import copy

from batch_shipyard_b200.backend.agent import NodeAgent
from batch_shipyard_b200.backend.local import LocalBackend
from batch_shipyard_b200.jobs import submit
from batch_shipyard_b200.pool import provision

BASE = {
    "credentials": {"storage": {"acct": {"account": "local"}}},
    "batch_shipyard": {"storage_account_settings": "acct", "store_timing_metrics": True},
    "global_resources": {"docker_images": ["busybox"]},
    "pool_specification": {"id": "testpool", "vm_size": "STANDARD_D2_V2", "vm_count": {"dedicated": 2, "low_priority": 0},
                           "inter_node_communication_enabled": True},
}


def make(tmp_path, tasks=None, job=None, pool=None, extra=None, jobs=None):
    cfg = copy.deepcopy(BASE)
    if pool:
        cfg["pool_specification"].update(pool)
    if jobs is not None:
        cfg["job_specifications"] = jobs
    else:
        j = {"id": "job1", "tasks": tasks or [{"docker_image": "busybox", "command": "echo hi"}]}
        j.update(job or {})
        cfg["job_specifications"] = [j]
    if extra:
        from batch_shipyard_b200.utils.util import merge_dict
        cfg = merge_dict(cfg, extra)
    b = LocalBackend(state_dir=str(tmp_path / "state"))
    return cfg, b


def up(cfg, b, **kw):
    return provision.create_pool(b, cfg, **kw)


def run(cfg, b, pool_id=None, max_seconds=60):
    out = submit.add_jobs(b, cfg)
    NodeAgent(b, pool_id or cfg["pool_specification"]["id"], poll=0.02).run(until_idle=True, max_seconds=max_seconds)
    return out


def read(b, job, task, name="stdout.txt"):
    with open(b.task_file_path(job, task, name)) as f:
        return f.read()
