"""Regression tests for the round-1 advisor findings: slot leaks on delete, adopted runners that vanish,
id validation, lease loss, encryption that must not degrade."""
import os
import signal
import subprocess
import time

import pytest

from _helpers import make, up
from batch_shipyard_b200 import crypto
from batch_shipyard_b200.backend.agent import NodeAgent, _AdoptedProc, LEASE_S
from batch_shipyard_b200.backend.local import BackendError, pid_alive, proc_start_ticks, validate_id
from batch_shipyard_b200.jobs import submit


def _drive(agent, until, timeout=30.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        agent.tick()
        if until():
            return True
        time.sleep(0.02)
    return False


def test_delete_running_job_frees_slot_and_kills_ranks(tmp_path):
    tasks = [{"id": "t", "docker_image": "busybox", "command": "echo $$ > $AZ_BATCH_NODE_SHARED_DIR/rank.pid; sleep 60"}]
    cfg, b = make(tmp_path, tasks=tasks, pool={"vm_count": {"dedicated": 1, "low_priority": 0}})
    up(cfg, b)
    import pathlib
    marker = pathlib.Path(b.node_shared_dir("testpool")) / "rank.pid"
    submit.add_jobs(b, cfg)
    agent = NodeAgent(b, "testpool", poll=0.02)
    assert agent.acquire()
    try:
        assert _drive(agent, lambda: marker.exists() and marker.read_text().strip())
        node = b.list_nodes("testpool")[0]
        assert node["state"] == "running" and node["running_tasks"] == [["job1", "t"]]
        rank_pid = int(marker.read_text())
        runner_pid = b.get_task("job1", "t")["pid"]
        b.delete_job("job1")
        assert not b.job_exists("job1")
        node = b.list_nodes("testpool")[0]
        assert node["state"] == "idle" and node["running_tasks"] == []       # slot back immediately, not "when the agent reaps"
        assert not pid_alive(rank_pid)                                        # rank torn down by the runner, not orphaned
        agent.tick()                                                          # reaping a runner whose rows are gone must not raise
        assert not pid_alive(runner_pid)
        # the node schedules again
        cfg["job_specifications"][0]["id"] = "job2"
        cfg["job_specifications"][0]["tasks"][0]["command"] = "echo again"
        submit.add_jobs(b, cfg)
        assert _drive(agent, lambda: b.get_task("job2", "t")["state"] == "completed")
        assert b.get_task("job2", "t")["result"] == "success"
    finally:
        agent.release()


def test_delete_running_task_and_stale_slot_sweep(tmp_path):
    tasks = [{"id": "t", "docker_image": "busybox", "command": "sleep 60"}]
    cfg, b = make(tmp_path, tasks=tasks, pool={"vm_count": {"dedicated": 1, "low_priority": 0}})
    up(cfg, b)
    submit.add_jobs(b, cfg)
    agent = NodeAgent(b, "testpool", poll=0.02)
    assert agent.acquire()
    try:
        assert _drive(agent, lambda: b.get_task("job1", "t")["state"] == "running")
        b.delete_task("job1", "t")
        assert b.list_nodes("testpool")[0]["running_tasks"] == []
        # a slot whose task row vanished behind the agent's back is swept by tick()
        def ghost(n):
            n["running_tasks"].append(["job1", "ghost"])
            n["state"] = "running"
        b.store.mutate("node", "testpool", "cpu-0", ghost)
        agent.tick()
        n = b.list_nodes("testpool")[0]
        assert n["running_tasks"] == [] and n["state"] == "idle"
    finally:
        agent.release()


def test_adopted_runner_that_vanishes_is_not_success(tmp_path):
    tasks = [{"id": "t", "docker_image": "busybox", "command": "sleep 60"}]
    cfg, b = make(tmp_path, tasks=tasks, pool={"vm_count": {"dedicated": 1, "low_priority": 0}})
    up(cfg, b)
    submit.add_jobs(b, cfg)
    a1 = NodeAgent(b, "testpool", poll=0.02)
    assert a1.acquire()
    assert _drive(a1, lambda: b.get_task("job1", "t")["state"] == "running")
    t = b.get_task("job1", "t")
    assert t["pid_start"] == proc_start_ticks(t["pid"])
    popen = a1.procs.pop(("job1", "t"))                    # the first agent "crashes": its runner lives on
    a1.release()
    a2 = NodeAgent(b, "testpool", poll=0.02)
    assert a2.acquire()
    try:
        a2.recover_orphans()
        assert isinstance(a2.procs[("job1", "t")], _AdoptedProc)
        os.kill(t["pid"], signal.SIGKILL)                   # dies without writing result.json
        popen.wait()
        assert a2.procs[("job1", "t")].poll() == _AdoptedProc.EXIT_UNKNOWN
        a2._reap()
        t2 = b.get_task("job1", "t")
        assert t2["state"] == "active" and t2.get("result") != "success" and t2["requeue_count"] == 1   # requeued, never "completed success 0"
        assert b.list_nodes("testpool")[0]["running_tasks"] == []
    finally:
        for p in a2.procs.values():
            try:
                os.kill(p.pid, signal.SIGTERM)
            except (OSError, AttributeError):
                pass
        a2.release()
    # pid reuse guard: a live pid with a different start time is not "our runner"
    assert pid_alive(os.getpid(), proc_start_ticks(os.getpid())) and not pid_alive(os.getpid(), 1)


@pytest.mark.parametrize("bad", ["..", "a/b", "x;rm -rf ~", "$(id)", "", "a" * 65, "sp ace"])
def test_ids_are_validated(tmp_path, bad):
    with pytest.raises(BackendError):
        validate_id("job", bad)
    cfg, b = make(tmp_path, job={"id": bad})
    up(cfg, b)
    with pytest.raises((submit.JobSubmissionError, BackendError, ValueError, KeyError)):
        submit.add_jobs(b, cfg)
    with pytest.raises(BackendError):
        b.add_job({"id": bad, "pool_id": "testpool"})
    with pytest.raises(BackendError):
        b.add_job_schedule({"id": bad, "pool_id": "testpool"})
    with pytest.raises(BackendError):
        b.delete_pool(bad)
    assert os.path.isdir(b.pool_root("testpool"))           # nothing outside was touched


def test_schedule_job_ids_still_accepted(tmp_path):
    cfg, b = make(tmp_path)
    up(cfg, b)
    b.add_job({"id": "sched1:job-3", "schedule_id": "sched1", "pool_id": "testpool"})
    with pytest.raises(BackendError):
        b.add_job({"id": "sched1:job-3/../x", "schedule_id": "sched1", "pool_id": "testpool"})


def test_agent_stops_when_lease_is_lost(tmp_path):
    cfg, b = make(tmp_path)
    up(cfg, b)
    agent = NodeAgent(b, "testpool", poll=0.02)
    assert agent.acquire()
    assert agent._ensure_lease()
    # lease expires and another agent takes it (what a > LEASE_S job-preparation command used to cause silently)
    b.store.release_lease(agent.lease_name, agent.holder)
    assert b.store.acquire_lease(agent.lease_name, "intruder", LEASE_S)
    assert not agent._ensure_lease() and agent.lease_lost
    agent2 = NodeAgent(b, "testpool", poll=0.02, holder=agent.holder)
    agent2.lease_lost = True
    with pytest.raises(BackendError):
        agent2.run(max_seconds=1)
    # expiry that nobody used is simply re-acquired
    b.store.release_lease(agent.lease_name, "intruder")
    agent3 = NodeAgent(b, "testpool", poll=0.02)
    assert agent3._ensure_lease()
    agent3.release()


def test_lease_keeper_renews_during_long_aux_command(tmp_path, monkeypatch):
    import batch_shipyard_b200.backend.agent as A
    monkeypatch.setattr(A, "LEASE_S", 0.6)
    cfg, b = make(tmp_path, job={"job_preparation": {"command": "sleep 1.5"}})
    up(cfg, b)
    submit.add_jobs(b, cfg)
    agent = NodeAgent(b, "testpool", poll=0.02)
    agent.run(until_idle=True, max_seconds=30)             # would lose a 0.6 s lease during the 1.5 s preparation without the keeper
    assert not agent.lease_lost
    assert b.get_task("job1", b.list_tasks("job1")[0]["id"])["result"] == "success"


def test_encryption_never_degrades(tmp_path):
    assert crypto.encrypt_string(False, "secret") == "secret"
    with pytest.raises(crypto.EncryptionError):
        crypto.encrypt_string(True, "secret", None)
    with pytest.raises(crypto.EncryptionError):
        crypto.encrypt_string(True, "secret", str(tmp_path / "missing.pem"))
    if subprocess.call(["which", "openssl"], stdout=subprocess.DEVNULL) != 0:
        pytest.skip("openssl not installed")
    out = crypto.generate_pem_pfx_certificates(str(tmp_path / "cert"), pfx_password="x")
    long_secret = "k" * 1000 + "é"                          # longer than one RSA block
    enc = crypto.encrypt_string(True, long_secret, out["pem"])
    assert enc.startswith("rsa:") and "," in enc and "kkkk" not in enc
    assert crypto.decrypt_string(enc, out["private_key"]) == long_secret
    with pytest.raises(crypto.EncryptionError):
        crypto.decrypt_string(enc, None)
    bad = tmp_path / "bad.pem"
    bad.write_text("not a certificate")
    with pytest.raises(crypto.EncryptionError):
        crypto.encrypt_string(True, "secret", str(bad))
