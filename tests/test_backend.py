"""Local backend + node agent + native runner: the integration layer of the test pyramid."""
import json
import os
import time

import pytest

from _helpers import make, read, run, up
from batch_shipyard_b200.backend.agent import NodeAgent
from batch_shipyard_b200.backend.local import BackendError
from batch_shipyard_b200.jobs import submit
from batch_shipyard_b200.pool import provision


def test_pool_lifecycle_and_markers(tmp_path):
    cfg, b = make(tmp_path)
    pool = up(cfg, b)
    assert pool["allocation_state"] == "steady" and pool["_summary"]["ready"] == 2
    nodes = b.list_nodes("testpool")
    assert [n["id"] for n in nodes] == ["cpu-0", "cpu-1"] and all(n["state"] == "idle" for n in nodes)
    assert os.path.exists(os.path.join(b.node_startup_dir("testpool"), "cpu-0", provision.NODEPREP_FINISHED))
    ev = [(e["source"], e["event"]) for e in b.store.events("testpool")]
    assert ("nodeprep", "start") in ev and ("nodeprep", "end") in ev and ("cascade", "start") in ev and ("cascade", "gr-done") in ev
    assert [e["event"] for e in b.store.events("testpool") if e["source"] == "cascade"].count("pull-end") == 1
    with pytest.raises(provision.PoolCreationError):
        up(cfg, b)
    up(cfg, b, recreate=True)
    b.resize_pool("testpool", 4, 0)
    provision.bring_up_nodes(b, "testpool")
    assert b.node_counts("testpool")["dedicated"]["idle"] == 4
    b.resize_pool("testpool", 1, 0)
    assert b.node_counts("testpool")["dedicated"]["total"] == 1
    stats = b.pool_stats("testpool")
    assert stats["total_nodes"] == 1 and stats["task_slots"] == 1
    b.delete_pool("testpool")
    assert not b.pool_exists("testpool")


def test_start_task_failure_reboot_and_unusable_recovery(tmp_path):
    cfg, b = make(tmp_path, pool={"reboot_on_start_task_failed": True, "attempt_recovery_on_unusable": True})
    calls = {}

    def hook(node, attempt):
        calls[node["id"]] = calls.get(node["id"], 0) + 1
        if node["id"] == "cpu-0" and attempt < 2:
            return "start task exited 1"
        if node["id"] == "cpu-1" and attempt == 0 and calls[node["id"]] == 1:
            return "unusable: device fell off the bus"
        return None
    pool = up(cfg, b, fault_hook=hook)
    s = pool["_summary"]
    assert s["rebooted"] == 2 and s["recovered"] == 1
    provision.bring_up_nodes(b, "testpool")
    states = {n["id"]: n["state"] for n in b.list_nodes("testpool")}
    assert states["cpu-0"] == "idle" and list(states.values()).count("idle") == 2
    # without reboot permission the node stays start_task_failed
    cfg2, b2 = make(tmp_path / "b")
    pool2 = up(cfg2, b2, fault_hook=lambda n, a: "boom" if n["id"] == "cpu-1" else None)
    assert pool2["_summary"]["start_task_failed"] == 1
    assert b2.list_nodes("testpool", start_task_failed=True)[0]["id"] == "cpu-1"


def test_cross_field_rules(tmp_path):
    cfg, _ = make(tmp_path, pool={"vm_count": {"dedicated": 1, "low_priority": 1}})
    with pytest.raises(ValueError):
        provision.adjust_settings_for_pool_creation(cfg)
    cfg, _ = make(tmp_path, pool={"per_job_auto_scratch": True, "inter_node_communication_enabled": False})
    with pytest.raises(ValueError):
        provision.adjust_settings_for_pool_creation(cfg)
    cfg, _ = make(tmp_path, pool={"transfer_files_on_pool_creation": True})
    warns = provision.adjust_settings_for_pool_creation(cfg)
    assert cfg["pool_specification"]["block_until_all_global_resources_loaded"] is False and warns
    assert "-b" not in provision.nodeprep_flags(cfg) and "-c" in provision.nodeprep_flags(cfg)


def test_runner_contract_env_and_exit_code(tmp_path):
    tasks = [{"id": "ok", "docker_image": "busybox", "environment_variables": {"FOO": "bar"},
              "command": 'echo "$AZ_BATCH_JOB_ID/$AZ_BATCH_TASK_ID $FOO $AZ_BATCH_POOL_ID"; test -d "$AZ_BATCH_TASK_WORKING_DIR" && test "$PWD" = "$AZ_BATCH_TASK_WORKING_DIR" && test -d "$AZ_BATCH_NODE_SHARED_DIR"'},
             {"id": "bad", "docker_image": "busybox", "command": "echo oops >&2; exit 7"}]
    cfg, b = make(tmp_path, tasks=tasks)
    up(cfg, b); run(cfg, b)
    assert read(b, "job1", "ok").strip() == "job1/ok bar testpool"
    t = b.get_task("job1", "ok")
    assert t["state"] == "completed" and t["result"] == "success" and t["exit_code"] == 0 and len(t["node_ids"]) == 1
    bad = b.get_task("job1", "bad")
    assert bad["result"] == "failure" and bad["exit_code"] == 7 and "oops" in read(b, "job1", "bad", "stderr.txt")
    envlist = read(b, "job1", "ok", ".shipyard.envlist")
    assert "FOO=bar" in envlist and "\nHOME=" not in "\n" + envlist and "\nPATH=" not in "\n" + envlist
    res = json.loads(read(b, "job1", "bad", "result.json"))
    assert res["exit_code"] == 7 and res["result"] == "fail"
    assert b.count_tasks("job1") == {"active": 0, "running": 0, "completed": 2, "succeeded": 1, "failed": 1}
    assert all(n["state"] == "idle" and not n["running_tasks"] for n in b.list_nodes("testpool"))


def test_dependencies_merge_task_and_ids(tmp_path):
    tasks = [{"docker_image": "busybox", "task_factory": {"repeat": 3}, "command": "echo part >> $AZ_BATCH_NODE_SHARED_DIR/parts"},
             {"id": "after", "docker_image": "busybox", "depends_on": ["task-00000", "task-00001"], "command": "wc -l < $AZ_BATCH_NODE_SHARED_DIR/parts"}]
    cfg, b = make(tmp_path, tasks=tasks, job={"merge_task": {"docker_image": "busybox", "command": "echo merged"}})
    up(cfg, b)
    out = run(cfg, b)
    assert out["job1"]["task_ids"] == ["task-00000", "task-00001", "task-00002", "after", "merge-task-00000"]
    m = b.get_task("job1", "merge-task-00000")
    assert m["result"] == "success" and sorted(m["depends_on"]) == ["after", "task-00000", "task-00001", "task-00002"]
    assert m["start_time"] >= max(b.get_task("job1", t)["end_time"] for t in m["depends_on"])
    assert int(read(b, "job1", "after").strip()) >= 2
    # appending to the existing job continues the generic id sequence
    with pytest.raises(submit.JobSubmissionError):
        submit.add_jobs(b, cfg)                       # explicit id "after" would collide
    cfg["job_specifications"][0]["tasks"].pop()
    out2 = submit.add_jobs(b, cfg)
    assert out2["job1"]["task_ids"][0] == "task-00003" and out2["job1"]["task_ids"][-1] == "merge-task-00001"


def test_dependency_failure_blocks_or_satisfies(tmp_path):
    def tasks(action):
        return [{"id": "a", "docker_image": "busybox", "command": "exit 3", "exit_conditions": {"default": {"exit_options": {"dependency_action": action}}}},
                {"id": "b", "docker_image": "busybox", "depends_on": ["a"], "command": "echo ran"}]
    cfg, b = make(tmp_path, tasks=tasks("block")); up(cfg, b); run(cfg, b, max_seconds=5)
    assert b.get_task("job1", "b")["state"] == "active"
    cfg, b = make(tmp_path / "s", tasks=tasks("satisfy")); up(cfg, b); run(cfg, b)
    assert b.get_task("job1", "b")["result"] == "success"
    cfg, b = make(tmp_path / "r", tasks=[{"id": "0", "docker_image": "busybox", "command": "true"}, {"id": "1", "docker_image": "busybox", "command": "true"},
                                         {"id": "r", "docker_image": "busybox", "depends_on_range": [0, 1], "command": "echo range"}])
    up(cfg, b); run(cfg, b)
    assert b.get_task("job1", "r")["result"] == "success"


def test_retries_exit_action_and_auto_complete(tmp_path):
    flaky = 'f=$AZ_BATCH_NODE_SHARED_DIR/flaky; n=$(cat $f 2>/dev/null || echo 0); echo $((n+1)) > $f; test $n -ge 2'
    cfg, b = make(tmp_path, tasks=[{"id": "flaky", "docker_image": "busybox", "max_task_retries": 3, "command": flaky}], job={"auto_complete": True})
    up(cfg, b); run(cfg, b)
    t = b.get_task("job1", "flaky")
    assert t["result"] == "success" and t["retry_count"] == 2
    assert b.get_job("job1")["state"] == "completed" and b.get_job("job1")["terminate_reason"] == "AllTasksComplete"
    tasks = [{"id": "boom", "docker_image": "busybox", "command": "exit 1", "exit_conditions": {"default": {"exit_options": {"job_action": "terminate"}}}},
             {"id": "later", "docker_image": "busybox", "depends_on": ["boom"], "command": "echo never"}]
    cfg, b = make(tmp_path / "t", tasks=tasks); up(cfg, b); run(cfg, b)
    assert b.get_job("job1")["state"] == "completed" and "boom" in b.get_job("job1")["terminate_reason"]
    assert b.get_task("job1", "later")["result"] == "failure"


def test_job_preparation_release_and_input_output_data(tmp_path):
    store = tmp_path / "acct"
    (store / "in" / "d").mkdir(parents=True)
    (store / "in" / "a.dat").write_text("A"); (store / "in" / "d" / "b.dat").write_text("B"); (store / "in" / "skip.tmp").write_text("x")
    tasks = [{"id": "t", "docker_image": "busybox",
              "input_data": {"azure_storage": [{"storage_account_settings": "acct", "remote_path": "in", "local_path": "$AZ_BATCH_TASK_WORKING_DIR/inp", "exclude": ["*.tmp"]}]},
              "output_data": {"azure_storage": [{"storage_account_settings": "acct", "remote_path": "out/ok", "local_path": "$AZ_BATCH_TASK_WORKING_DIR/res", "condition": "tasksuccess"},
                                                {"storage_account_settings": "acct", "remote_path": "out/fail", "local_path": "$AZ_BATCH_TASK_WORKING_DIR/res", "condition": "taskfailure"}]},
              "command": "mkdir res; cat inp/a.dat inp/d/b.dat > res/joined; test ! -e inp/skip.tmp; cat $AZ_BATCH_NODE_SHARED_DIR/prep"}]
    cfg, b = make(tmp_path, tasks=tasks, job={"auto_complete": True, "job_preparation": {"command": "echo prepared > $AZ_BATCH_NODE_SHARED_DIR/prep"},
                                               "job_release": {"command": "echo released > $AZ_BATCH_NODE_SHARED_DIR/rel"}},
                  extra={"credentials": {"storage": {"acct": {"local_path": str(store)}}}})
    up(cfg, b); run(cfg, b)
    assert b.get_task("job1", "t")["result"] == "success", read(b, "job1", "t", "stderr.txt")
    assert read(b, "job1", "t").strip() == "prepared"
    assert (store / "out" / "ok" / "joined").read_text() == "AB" and not (store / "out" / "fail").exists()
    assert open(os.path.join(b.node_shared_dir("testpool"), "rel")).read().strip() == "released"
    assert b.get_job("job1")["state"] == "completed"


def test_multi_instance_rank_env_and_failure_propagation(tmp_path):
    mi = {"num_instances": "pool_current_dedicated", "coordination_command": "echo coord >> $AZ_BATCH_NODE_SHARED_DIR/coord",
          "pre_execution_command": "echo pre > $AZ_BATCH_NODE_SHARED_DIR/pre", "mpi": {"runtime": "openmpi", "processes_per_node": 2}}
    tasks = [{"id": "mi", "docker_image": "busybox", "multi_instance": mi,
              "command": 'echo "rank $RANK/$WORLD_SIZE ompi $OMPI_COMM_WORLD_RANK master $AZ_BATCH_IS_CURRENT_NODE_MASTER hosts $AZ_BATCH_HOST_LIST" > $AZ_BATCH_TASK_DIR/r$RANK.txt'}]
    cfg, b = make(tmp_path, tasks=tasks); up(cfg, b); run(cfg, b)
    t = b.get_task("job1", "mi")
    assert t["result"] == "success" and sorted(t["node_ids"]) == ["cpu-0", "cpu-1"]
    assert t["mpi_command"].startswith("mpirun --oversubscribe -host $AZ_BATCH_HOST_LIST -np 4 --map-by ppr:2:node")
    got = sorted(read(b, "job1", "mi", f"r{r}.txt").strip() for r in range(4))
    assert got[0] == "rank 0/4 ompi 0 master true hosts 127.0.0.1,127.0.0.1" and got[3].startswith("rank 3/4 ompi 3 master false")
    assert open(os.path.join(b.node_shared_dir("testpool"), "coord")).read().count("coord") == 2
    # a dying rank fails the whole task and the survivors are reaped (no hang)
    tasks = [{"id": "die", "docker_image": "busybox", "multi_instance": {"num_instances": 2, "mpi": {"runtime": "mpich", "processes_per_node": 1}},
              "command": 'if [ "$RANK" = "1" ]; then exit 9; else sleep 60; fi'}]
    cfg, b = make(tmp_path / "d", tasks=tasks); up(cfg, b)
    t0 = time.time(); run(cfg, b)
    d = b.get_task("job1", "die")
    assert d["result"] == "failure" and d["exit_code"] == 9 and time.time() - t0 < 30
    assert d["failure_info"]["rank_exit_codes"][1] == 9
    # a task needing more instances than the pool has is never placed
    cfg, b = make(tmp_path / "n", tasks=[{"id": "big", "docker_image": "busybox", "multi_instance": {"num_instances": 3}, "command": "true"}])
    up(cfg, b); run(cfg, b, max_seconds=3)
    assert b.get_task("job1", "big")["state"] == "active" and "needs 3 instances" in b.get_task("job1", "big")["scheduling_note"]


def test_wall_time_terminate_and_fault_injection(tmp_path, monkeypatch):
    cfg, b = make(tmp_path, tasks=[{"id": "slow", "docker_image": "busybox", "max_wall_time": "00:00:01", "command": "sleep 30"}])
    up(cfg, b); t0 = time.time(); run(cfg, b)
    t = b.get_task("job1", "slow")
    assert t["result"] == "failure" and t["exit_code"] == 124 and time.time() - t0 < 20 and "wall time" in t["failure_info"]["message"]
    cfg, b = make(tmp_path / "k", tasks=[{"id": "victim", "docker_image": "busybox", "command": "sleep 30"}])
    up(cfg, b); submit.add_jobs(b, cfg)
    agent = NodeAgent(b, "testpool", poll=0.02)
    assert agent.acquire()
    agent.tick()
    assert b.get_task("job1", "victim")["state"] == "running"
    b.terminate_task("job1", "victim")
    for _ in range(400):
        agent.tick()
        if b.get_task("job1", "victim")["state"] == "completed":
            break
        time.sleep(0.02)
    v = b.get_task("job1", "victim")
    assert v["state"] == "completed" and v["result"] == "failure" and "terminated" in v["failure_info"]["message"]
    agent.release()
    monkeypatch.setenv("SHIPYARD_FAULT_INJECT", "kill_rank:0:after_ms:100")
    cfg, b = make(tmp_path / "f", tasks=[{"id": "inj", "docker_image": "busybox", "multi_instance": {"num_instances": 2, "mpi": {"runtime": "openmpi", "processes_per_node": 1}}, "command": "sleep 20"}])
    up(cfg, b); t0 = time.time(); run(cfg, b)
    assert b.get_task("job1", "inj")["exit_code"] == 137 and time.time() - t0 < 15


def test_disable_enable_migrate_and_priority(tmp_path):
    cfg, b = make(tmp_path, tasks=[{"id": "t", "docker_image": "busybox", "command": "echo $AZ_BATCH_POOL_ID"}])
    up(cfg, b)
    cfg2, _ = make(tmp_path, pool={"id": "other", "vm_count": {"dedicated": 1, "low_priority": 0}})
    up(cfg2, b)
    submit.add_jobs(b, cfg)
    b.disable_job("job1", "requeue")
    NodeAgent(b, "testpool", poll=0.02).run(until_idle=True, max_seconds=2)
    assert b.get_task("job1", "t")["state"] == "active"
    with pytest.raises(BackendError):
        b.migrate_job("nope", "other")
    b.migrate_job("job1", "other"); b.enable_job("job1")
    NodeAgent(b, "other", poll=0.02).run(until_idle=True, max_seconds=30)
    assert read(b, "job1", "t").strip() == "other"
    st = b.job_stats()
    assert st["jobs"] == 1 and st["succeeded"] == 1


def test_recurrence_job_schedule(tmp_path):
    cfg, b = make(tmp_path, tasks=[{"docker_image": "busybox", "command": "date +%s%N >> $AZ_BATCH_NODE_SHARED_DIR/ticks"}],
                  job={"auto_complete": True, "recurrence": {"schedule": {"recurrence_interval": "00:01:00"}}})
    up(cfg, b)
    out = submit.add_jobs(b, cfg)
    assert out["job1"]["kind"] == "job_schedule" and out["job1"]["tasks_per_recurrence"] == 1
    b.store.merge("jobschedule", "job1", "", {"recurrence_interval_s": 0.3})      # speed the clock up for the test
    NodeAgent(b, "testpool", poll=0.02).run(until_idle=False, max_seconds=2.0)
    s = b.get_job_schedule("job1")
    assert s["runs"] >= 3
    ticks = open(os.path.join(b.node_shared_dir("testpool"), "ticks")).read().split()
    assert len(ticks) >= 3 and b.job_exists("job1:job-1")
    b.terminate_job_schedule("job1")
    assert b.get_job_schedule("job1")["state"] == "completed"


def test_agent_crash_recovery_and_lease(tmp_path):
    cfg, b = make(tmp_path, tasks=[{"id": "t", "docker_image": "busybox", "command": "echo again"}])
    up(cfg, b); submit.add_jobs(b, cfg)
    # simulate an agent that died right after marking the task running
    b.update_task("job1", "t", state="running", pid=2 ** 22 + 12345, node_ids=["cpu-0"])
    b.store.mutate("node", "testpool", "cpu-0", lambda n: (n["running_tasks"].append(["job1", "t"]), n.update(state="running")) and None)
    a1 = NodeAgent(b, "testpool", poll=0.02)
    assert a1.acquire()
    with pytest.raises(BackendError):
        NodeAgent(b, "testpool").run()           # lease is held
    a1.release()
    NodeAgent(b, "testpool", poll=0.02).run(until_idle=True, max_seconds=30)
    t = b.get_task("job1", "t")
    assert t["result"] == "success" and t["requeue_count"] == 1


def test_auto_scratch_and_missing_image_policy(tmp_path):
    cfg, b = make(tmp_path, pool={"per_job_auto_scratch": True},
                  tasks=[{"id": "w", "docker_image": "busybox", "command": "touch $SHIPYARD_AUTO_SCRATCH/x && ls $SHIPYARD_AUTO_SCRATCH"}],
                  job={"auto_scratch": {"setup": "dependency", "num_instances": "pool_current_dedicated"}, "auto_complete": True})
    up(cfg, b); out = run(cfg, b)
    assert out["job1"]["task_ids"][0] == "batch-shipyard-autoscratch"
    assert b.get_task("job1", "w")["result"] == "success" and read(b, "job1", "w").strip() == "x"
    assert "batch-shipyard-autoscratch" in b.get_task("job1", "w")["depends_on"]
    assert not os.path.exists(os.path.join(b.node_shared_dir("testpool"), "auto_scratch", "job1"))   # job release cleaned it
    cfg, b = make(tmp_path / "m", tasks=[{"docker_image": "notlisted", "command": "true"}]); up(cfg, b)
    with pytest.raises(submit.JobSubmissionError):
        submit.add_jobs(b, cfg)
    cfg["job_specifications"][0]["allow_run_on_missing_image"] = True
    submit.add_jobs(b, cfg)


def test_max_tasks_per_node_and_fill(tmp_path):
    cfg, b = make(tmp_path, pool={"vm_count": {"dedicated": 2, "low_priority": 0}, "max_tasks_per_node": 2, "node_fill_type": "pack"},
                  tasks=[{"docker_image": "busybox", "task_factory": {"repeat": 2}, "command": "sleep 0.4; echo $AZ_BATCH_NODE_ID"}])
    up(cfg, b); run(cfg, b)
    nodes = {read(b, "job1", t["id"]).strip() for t in b.list_tasks("job1")}
    assert nodes == {"cpu-0"}                                  # packed onto one node
    cfg, b = make(tmp_path / "s", pool={"max_tasks_per_node": 2, "node_fill_type": "spread"},
                  tasks=[{"docker_image": "busybox", "task_factory": {"repeat": 2}, "command": "sleep 0.4; echo $AZ_BATCH_NODE_ID"}])
    up(cfg, b); run(cfg, b)
    assert {read(b, "job1", t["id"]).strip() for t in b.list_tasks("job1")} == {"cpu-0", "cpu-1"}


def test_taskrun_accepts_the_reference_runner_environment_contract(tmp_path):
    """`shipyard-taskrun` without --spec behaves like the reference's shipyard_task_runner.sh: prologues, env file minus exclusions,
    `$RUNTIME $CMD $OPTS $IMAGE $USER_CMD`, epilogue with SHIPYARD_TASK_RESULT, exit code of the task."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "batch_shipyard_b200", "_native", "shipyard-taskrun")
    if not os.path.exists(exe):
        pytest.skip("native runner not built")
    envfile = tmp_path / "envlist"
    env = {"PATH": os.environ["PATH"], "KEEP": "y", "SECRET": "x", "SHIPYARD_ENV_EXCLUDE": "^SECRET=", "SHIPYARD_ENV_FILE": str(envfile),
           "SHIPYARD_SYSTEM_PROLOGUE_CMD": "echo sys-pro", "SHIPYARD_USER_PROLOGUE_CMD": "echo user-pro",
           "SHIPYARD_USER_CMD": "echo hello; exit 3", "SHIPYARD_SYSTEM_EPILOGUE_CMD": "echo result=$SHIPYARD_TASK_RESULT"}
    p = subprocess.run([exe], env=env, stdout=subprocess.PIPE, text=True, cwd=str(tmp_path))
    assert p.returncode == 3 and p.stdout.split() == ["sys-pro", "user-pro", "hello", "result=fail"]
    lines = envfile.read_text().splitlines()
    assert "KEEP=y" in lines and not any(ln.startswith("SECRET=") for ln in lines)
    # container form: runtime + command + expanded options + image + user command
    env2 = {"PATH": os.environ["PATH"], "SHIPYARD_RUNTIME": "echo", "SHIPYARD_RUNTIME_CMD": "run", "SHIPYARD_RUNTIME_CMD_OPTS": "--rm -e HOME=$HOME",
            "HOME": "/h", "SHIPYARD_CONTAINER_IMAGE_NAME": "busybox", "SHIPYARD_USER_CMD": "true", "SHIPYARD_SYSTEM_EPILOGUE_CMD": "echo $SHIPYARD_TASK_RESULT"}
    p = subprocess.run([exe], env=env2, stdout=subprocess.PIPE, text=True, cwd=str(tmp_path))
    assert p.returncode == 0 and p.stdout.splitlines() == ["run --rm -e HOME=/h busybox true", "success"]
    # a failing prologue aborts with its exit code (the reference script runs under `set -e`)
    p = subprocess.run([exe], env={"PATH": os.environ["PATH"], "SHIPYARD_SYSTEM_PROLOGUE_CMD": "exit 7", "SHIPYARD_USER_CMD": "echo no"},
                       stdout=subprocess.PIPE, text=True, cwd=str(tmp_path))
    assert p.returncode == 7 and "no" not in p.stdout


def test_large_sweep_submission_is_linear(tmp_path):
    """4000 generated tasks: ids are contiguous, submission is one transaction per collection of 100 and does not rescan the job per
    task (was quadratic: 10 s for 5000 tasks, now well under a second)."""
    import time
    from batch_shipyard_b200.jobs import builder as B
    cfg, b = make(tmp_path, tasks=[{"docker_image": "busybox", "command": "echo {0}",
                                    "task_factory": {"parametric_sweep": {"product": [{"start": 0, "stop": 4000, "step": 1}]}}}])
    up(cfg, b)
    t0 = time.time()
    out = submit.add_jobs(b, cfg)
    dt = time.time() - t0
    tasks = b.list_tasks("job1")
    assert len(tasks) == 4000 and out["job1"]["num_tasks"] == 4000
    assert sorted(t["id"] for t in tasks) == [f"task-{i:05d}" for i in range(4000)]
    assert dt < 8.0, dt
    # allocator semantics = the one-shot function applied repeatedly: continues after the highest id, explicit ids move the counter
    a = B.TaskIdAllocator({"task-00003", "other"}, set())
    got = [a.next("task-", 5), a.next("task-", 5)]
    a.reserve("task-00010")
    got += [a.next("task-", 5), a.next("task-", 5, is_merge=True)]
    assert got == ["task-00004", "task-00005", "task-00011", "merge-task-00000"]
    with pytest.raises(Exception):
        b.add_tasks("job1", [{"id": "task-00001"}])                     # a clash rolls the whole collection back


def test_job_preparation_runs_once_per_node_even_when_many_tasks_start_in_one_pass(tmp_path):
    """Regression: the scheduler reused a stale job snapshot inside one scheduling pass, so every task launched in that pass re-ran
    the job preparation on its node (80 preparations for 100 tasks on 4 nodes)."""
    cfg, b = make(tmp_path, pool={"vm_count": {"dedicated": 3, "low_priority": 0}, "max_tasks_per_node": 4},
                  job={"job_preparation": {"command": "echo prep-$AZ_BATCH_NODE_ID >> $AZ_BATCH_NODE_SHARED_DIR/prep.log"}},
                  tasks=[{"docker_image": "busybox", "command": "true", "task_factory": {"repeat": 40}}])
    up(cfg, b)
    run(cfg, b, max_seconds=120)
    tasks = b.list_tasks("job1")
    assert len(tasks) == 40 and all(t["state"] == "completed" and t["result"] == "success" for t in tasks)
    lines = open(os.path.join(b.node_shared_dir("testpool"), "prep.log")).read().split()
    assert len(lines) == len(set(lines)) <= 3 and len(lines) >= 1          # at most once per node, never twice on the same node
    assert sorted(b.get_job("job1")["prep_nodes"]) == sorted(ln[len("prep-"):] for ln in lines)


@pytest.mark.parametrize("fill,expect", [("pack", [4, 0]), ("spread", [2, 2])])
def test_node_fill_type_places_tasks_pack_or_spread(tmp_path, fill, expect):
    """Four tasks that start in ONE scheduling pass on a 2-node x 4-slot pool: `pack` fills a node first, `spread` alternates
    (the per-pass node snapshot must be refreshed after every launch for this to hold)."""
    from batch_shipyard_b200.backend.agent import NodeAgent
    cfg, b = make(tmp_path, pool={"vm_count": {"dedicated": 2, "low_priority": 0}, "max_tasks_per_node": 4, "node_fill_type": fill},
                  tasks=[{"docker_image": "busybox", "command": "sleep 1.5", "task_factory": {"repeat": 4}}])
    up(cfg, b)
    submit.add_jobs(b, cfg)
    agent = NodeAgent(b, "testpool", poll=0.02)
    agent.tick()                                                       # one pass launches all four
    per_node = sorted((len(n["running_tasks"]) for n in b.list_nodes("testpool")), reverse=True)
    assert per_node == expect, per_node
    agent.run(until_idle=True, max_seconds=60)
    assert all(t["result"] == "success" for t in b.list_tasks("job1"))


def test_rendezvous_session_is_unique_per_installation_and_survives_long_ids(tmp_path):
    """Two state directories running identically named pool / job / task must not share a rendezvous name (abstract socket, POSIX shm are
    machine-global), and 64-character ids must not push the distinguishing tail past the 106 characters the socket name keeps."""
    import re
    from batch_shipyard_b200.backend import runspec
    from batch_shipyard_b200.backend.local import LocalBackend

    def session_of(state, pid, jid, tid, retry=0):
        b = LocalBackend(state_dir=str(state))
        t = {"id": tid, "command": "true", "retry_count": retry}
        spec, _ = runspec.build_task_spec(b, {"id": pid}, {"id": jid}, t, [{"id": "n0", "gpu_index": None}])
        return re.search(r"^session\t(.*)$", open(spec).read(), re.M).group(1)

    a = session_of(tmp_path / "a", "p", "j", "t")
    assert a != session_of(tmp_path / "b", "p", "j", "t") and a == session_of(tmp_path / "a", "p", "j", "t")
    long_ids = ("p" * 64, "j" * 64, "t" * 64)
    s0, s1 = session_of(tmp_path / "a", *long_ids, retry=0), session_of(tmp_path / "a", *long_ids, retry=1)
    assert s0 != s1 and len(s0) <= 72 and len(s1) <= 72 and s0[:40] == s1[:40]
