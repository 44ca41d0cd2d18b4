"""Federation constraints / best fit / daemon, and the CLI surface."""
import json
import os

def _read(path):
    with open(path) as f:
        return f.read()


import pytest
from click.testing import CliRunner

from _helpers import make, read, up
from batch_shipyard_b200 import cli
from batch_shipyard_b200.backend.agent import NodeAgent
from batch_shipyard_b200.fed import client as FCl
from batch_shipyard_b200.fed import constraints as FC
from batch_shipyard_b200.fed.daemon import FederationProcessor
from batch_shipyard_b200.fed.scheduler import select_pool
from batch_shipyard_b200.jobs import submit


def view(**kw):
    d = dict(id="p", valid=True, location="local", vm_size="b200x8", native=False, windows=False, autoscale_enabled=False,
             target_low_priority=0, max_tasks_per_node=1, inter_node_communication=True, cores_per_node=8, memory_mb_per_node=65536.0,
             registries=[], idle_dedicated=2, schedulable_dedicated=2)
    d.update(kw)
    return FC.PoolView(**d)


def cons(fc=None, tasks=None, **job):
    return FC.parse_constraints(dict({"federation_constraints": fc or {}}, **job), tasks or [{"id": "t"}])


def test_hard_constraint_truth_table():
    ok = view()
    assert FC.first_failed_hard_constraint(ok, cons()) is None
    cases = [
        ({"pool": {"location": "eastus"}}, view(), "location"),
        ({"pool": {"native": True}}, view(), "native"),
        ({"pool": {"windows": True}}, view(), "windows"),
        ({"pool": {"autoscale": {"allow": False}}}, view(autoscale_enabled=True), "autoscale_allow"),
        ({"pool": {"autoscale": {"allow": True, "exclusive": True}}}, view(), "autoscale_exclusive"),
        ({"pool": {"low_priority_nodes": {"allow": False}}}, view(target_low_priority=2), "low_priority_nodes_allow"),
        ({"pool": {"low_priority_nodes": {"allow": True, "exclusive": True}}}, view(), "low_priority_nodes_exclusive"),
        ({"compute_node": {"exclusive": True}}, view(max_tasks_per_node=4), "exclusive"),
        ({"compute_node": {"vm_size": "STANDARD_F1"}}, view(), "vm_size"),
        ({"compute_node": {"gpu": False}}, view(), "gpu"),
        ({"compute_node": {"infiniband": False}}, view(), "infiniband"),
        ({"compute_node": {"cores": {"amount": 16}}}, view(), "cores"),
        ({"compute_node": {"cores": {"amount": 4, "schedulable_variance": 0}}}, view(), "cores"),
        ({"compute_node": {"cores": {"amount": 4, "schedulable_variance": 0.5}}}, view(), "cores"),
        ({"compute_node": {"memory": {"amount": "128g"}}}, view(), "memory"),
        ({"pool": {"container_registries": {"public": ["my.reg.io"]}}}, view(), "registries"),
    ]
    for fc, p, name in cases:
        got = FC.first_failed_hard_constraint(p, cons(fc))
        assert got is not None and got[0] == name, (fc, got)
    assert FC.first_failed_hard_constraint(view(valid=False), cons())[0] == "valid"
    assert FC.first_failed_hard_constraint(view(), cons({"compute_node": {"cores": {"amount": 4, "schedulable_variance": 1.0}}})) is None
    mi = cons(tasks=[{"id": "t", "multi_instance": {"num_instances": 4}}])
    assert mi.task.has_multi_instance and mi.task.instance_counts_max == 4
    assert FC.first_failed_hard_constraint(view(inter_node_communication=False), mi)[0] == "has_multi_instance"
    backlog = cons({"pool": {"max_active_task_backlog": {"ratio": 0.5, "autoscale_exempt": True}}})
    assert FC.fails_node_constraints(view(active_tasks=3), backlog)[0] == "max_active_task_backlog"
    assert FC.fails_node_constraints(view(active_tasks=3, autoscale_enabled=True), backlog) is None
    assert cons({"compute_node": {"memory": {"amount": "512m"}}}).compute_node.memory == 512
    with pytest.raises(ValueError):
        cons({"pool": {"autoscale": {"allow": False, "exclusive": True}}})


def test_greedy_best_fit_order():
    small = view(id="small", idle_dedicated=2, schedulable_dedicated=2)
    big = view(id="big", idle_dedicated=8, schedulable_dedicated=8)
    busy = view(id="busy", idle_dedicated=0, schedulable_dedicated=4, active_tasks=9)
    one = cons(tasks=[{"id": "a"}, {"id": "b"}])
    assert select_pool([big, small, busy], one)[0] == "small"                     # tightest idle fit
    four = cons(tasks=[{"id": str(i)} for i in range(4)])
    assert select_pool([small, busy, big], four)[0] == "big"
    assert select_pool([small, busy], four)[0] == "busy"                           # available (not idle) capacity
    mi8 = cons(tasks=[{"id": "m", "multi_instance": {"num_instances": 8}}])
    assert select_pool([small, busy], mi8)[0] is None                              # MI matched by nodes, never by backlog
    assert select_pool([small, busy, big], mi8)[0] == "big"
    auto = view(id="auto", idle_dedicated=0, schedulable_dedicated=0, autoscale_enabled=True)
    assert select_pool([auto], four)[0] == "auto"
    nine = cons(tasks=[{"id": str(i)} for i in range(9)])
    pid, diag = select_pool([small, busy], nine)
    assert pid == "small" and "backlog" in diag["small"]
    pid, diag = select_pool([view(id="bo", blackout_until=2e9)], one, now=1e9)
    assert pid is None and "blackout" in diag["bo"]


def test_federation_end_to_end(tmp_path, monkeypatch):
    monkeypatch.setenv("SHIPYARD_FED_NO_AGENT", "1")
    cfg, b = make(tmp_path, tasks=[{"docker_image": "busybox", "command": "echo $AZ_BATCH_POOL_ID"}], job={"auto_complete": True})
    up(cfg, b)
    cfg2, _ = make(tmp_path, pool={"id": "gpuish", "vm_size": "B200x8", "gpu": {"ignore_warnings": True}, "vm_count": {"dedicated": 1, "low_priority": 0}})
    up(cfg2, b)
    FCl.create_federation(b, "Fed1")
    with pytest.raises(FCl.FederationError):
        FCl.create_federation(b, "fed1")
    FCl.add_pools(b, "fed1", ["testpool", "gpuish"])
    cfg["job_specifications"][0]["federation_constraints"] = {"compute_node": {"gpu": True}}
    info = submit.add_jobs(b, cfg, federation_id="fed1")["job1"]
    assert info["kind"] == "job" and info["num_tasks"] == 1 and info["federation"]["id"] == "fed1"
    assert len(FCl.list_jobs(b, "fed1", queued=True)["queued"]) == 1
    with pytest.raises(FCl.FederationError):
        submit.add_jobs(b, cfg, federation_id="fed1") if b.store.exists("fedjob", "fed1", "job1") else (_ for _ in ()).throw(FCl.FederationError("x"))
    proc = FederationProcessor(b, blackout=0.0, evaluate_autoscale=False)
    assert proc.is_leader() and not FederationProcessor(b).is_leader()            # single leader via the lease
    assert proc.process_all() == 1
    fj = FCl.list_jobs(b, "fed1")["jobs"]["job1"]
    assert fj["pool_id"] == "gpuish"                                              # the gpu constraint steered it
    NodeAgent(b, "gpuish", poll=0.02).run(until_idle=True, max_seconds=30)
    assert read(b, "job1", "task-00000").strip() == "gpuish"
    # unschedulable -> blocked, not dropped; zap removes it
    cfg["job_specifications"][0]["id"] = "job2"
    cfg["job_specifications"][0]["federation_constraints"] = {"pool": {"location": "mars"}}
    uid = submit.add_jobs(b, cfg, federation_id="fed1")["job2"]["unique_id"]
    assert proc.process_all() == 0
    blk = FCl.list_jobs(b, "fed1", blocked=True)["blocked"]
    assert blk and blk[0]["unique_id"] == uid and "location" in blk[0]["reason"]
    assert FCl.zap_action(b, "fed1", uid)["removed"] >= 1
    assert not FCl.list_jobs(b, "fed1", blocked=True)["blocked"]
    # terminate through the action queue keeps FIFO per job
    FCl.enqueue_job_action(b, "fed1", "delete", ["job1"])
    assert proc.process_all() == 1 and not b.job_exists("job1")
    FCl.destroy_federation(b, "fed1")
    assert FCl.list_federations(b) == {}


def test_cli_surface_and_commands(tmp_path, monkeypatch):
    leaves = cli.leaf_commands()
    assert len(leaves) == 105
    for must in ("pool add", "pool autoscale evaluate", "pool nodes zap", "jobs tasks list", "data files stream", "fed jobs zap",
                 "fs cluster orchestrate", "monitor destroy", "slurm ssh node", "storage sas create", "keyvault add", "cert create",
                 "misc mirror-images", "diag logs upload", "account quota"):
        assert must in leaves
    assert sorted(cli.cli.commands) == ["account", "cert", "data", "diag", "fed", "fs", "jobs", "keyvault", "misc", "monitor", "pool", "slurm", "storage"]
    recipe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "recipes", "mpiBench-OpenMPI", "config")
    monkeypatch.setenv("SHIPYARD_STATE_DIR", str(tmp_path / "state"))
    monkeypatch.setenv("SHIPYARD_INLINE_AGENT", "1")
    r = CliRunner()
    res = r.invoke(cli.cli, ["pool", "exists", "--configdir", recipe], obj=cli.CliContext())
    assert res.exit_code == 1 and "does not exist" in res.output
    res = r.invoke(cli.cli, ["jobs", "add", "--configdir", recipe, "--dry-run", "--raw"], obj=cli.CliContext())
    assert res.exit_code == 0, res.output
    dry = json.loads(res.output)["mpibench"]
    assert dry["dry_run"] and dry["tasks"][0]["mpi_command"].startswith("mpirun --oversubscribe -host $AZ_BATCH_HOST_LIST -np 2")
    assert dry["tasks"][0]["multi_instance"]["coordination_line"].startswith("docker run -d")
    res = r.invoke(cli.cli, ["pool", "add", "--configdir", recipe, "--raw", "-y"], obj=cli.CliContext())
    assert res.exit_code == 0, res.output
    assert json.loads(res.output)["mpibench"]["node_counts"]["dedicated"]["idle"] == 2
    assert r.invoke(cli.cli, ["pool", "exists", "--configdir", recipe], obj=cli.CliContext()).exit_code == 0
    res = r.invoke(cli.cli, ["jobs", "add", "--configdir", recipe, "--raw"], obj=cli.CliContext())
    assert res.exit_code == 0, res.output
    res = r.invoke(cli.cli, ["jobs", "tasks", "list", "--configdir", recipe, "--raw"], obj=cli.CliContext())
    tasks = json.loads(res.output)["mpibench"]
    assert tasks[0]["state"] == "completed" and tasks[0]["result"] == "success" and tasks[0]["multi_instance"]
    res = r.invoke(cli.cli, ["data", "files", "stream", "--filespec", "mpibench,task-00000,stdout.txt", "--configdir", recipe], obj=cli.CliContext())
    assert "Allreduce" in res.output and "check failures: 0" in res.output
    for args in (["pool", "stats"], ["pool", "nodes", "list"], ["pool", "nodes", "count"], ["pool", "images", "list"], ["jobs", "list"],
                 ["jobs", "stats"], ["jobs", "tasks", "count"], ["account", "info"], ["account", "quota"], ["pool", "list"],
                 ["misc", "mirror-images"], ["pool", "rdp"], ["storage", "sas", "create", "acct", "cont/x"], ["monitor", "list"]):
        res = r.invoke(cli.cli, args + ["--configdir", recipe, "--raw"], obj=cli.CliContext())
        assert res.exit_code == 0, (args, res.output)
        json.loads(res.output)
    bad = tmp_path / "bad"
    bad.mkdir()
    (bad / "pool.yaml").write_text("pool_specification:\n  id: x\n  nonsense: 1\n")
    res = r.invoke(cli.cli, ["pool", "add", "--configdir", str(bad)], obj=cli.CliContext())
    assert res.exit_code == 1 and "unknown key" in res.output
    res = r.invoke(cli.cli, ["pool", "del", "--configdir", recipe, "-y", "--raw"], obj=cli.CliContext())
    assert json.loads(res.output)["deleted"] is True


def test_slurm_hostlist_and_retry_daemon(tmp_path):
    """slurmctld-facing entry points: hostlist expansion, resume failure -> retry queue -> daemon pass."""
    from batch_shipyard_b200.backend.local import LocalBackend
    from batch_shipyard_b200.slurm import cluster as sl
    assert sl.expand_hostlist("c-p-pool-[0-2,5],login0") == ["c-p-pool-0", "c-p-pool-1", "c-p-pool-2", "c-p-pool-5", "login0"]
    assert sl.expand_hostlist("n[08-10]") == ["n08", "n09", "n10"]
    b = LocalBackend(state_dir=str(tmp_path / "st"))
    cfg = {"slurm": {"cluster_id": "sc", "slurm_options": {"elastic_partitions": {}}}}
    cid = "sc"
    b.store.insert("slurmhost", cid, "sc-p-nopool-0", {"state": "suspended", "pool_id": "missing-pool", "resume_failures": 0})
    r = sl.resume(b, cfg, ["sc-p-nopool-0", "ghost"])
    assert [f["host"] for f in r["failed"]] == ["sc-p-nopool-0", "ghost"]
    assert b.store.get("slurmhost", cid, "sc-p-nopool-0")["resume_failures"] == 1
    assert len(b.store.peek_messages(f"slurm-retry-{cid}")) == 1
    out = sl.daemon(b, cfg, poll_interval=0.0, max_iterations=1)       # the retry fails again -> a new retry message, count 2
    assert out["iterations"] == 1 and out["dropped"] == 1
    assert b.store.get("slurmhost", cid, "sc-p-nopool-0")["resume_failures"] == 2
    sl.resume_failed(b, cfg, ["sc-p-nopool-0"])
    assert b.store.get("slurmhost", cid, "sc-p-nopool-0")["state"] == "suspended"


def test_federation_leader_failover(tmp_path, monkeypatch):
    """Two proxy daemons share the store: one leads; when it stops renewing, the lease expires and the other takes over."""
    import time
    from batch_shipyard_b200.backend.local import LocalBackend
    from batch_shipyard_b200.fed import daemon as fd
    monkeypatch.setattr(fd, "LEADER_LEASE_S", 0.3)
    monkeypatch.setattr(fd, "LEADER_RENEW_S", 0.05)
    b = LocalBackend(state_dir=str(tmp_path / "st"))
    a, c = fd.FederationProcessor(b, holder="proxy-a"), fd.FederationProcessor(b, holder="proxy-b")
    assert a.is_leader() and not c.is_leader()
    for _ in range(4):                        # a keeps renewing: b never gets in
        time.sleep(0.1)
        assert a.is_leader() and not c.is_leader()
    time.sleep(0.45)                          # a "crashes" (stops renewing): the lease runs out
    assert c.is_leader()
    assert not a.is_leader()                  # and the old leader cannot reclaim it while b renews
    assert b.store.lease_holder("federation-leader") == "proxy-b"


def test_generated_docs_are_current():
    """docs/cli.md lists every leaf command; regenerate with `python docs/gen_docs.py` after changing the CLI."""
    import os
    from batch_shipyard_b200 import cli
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = _read(os.path.join(root, "docs", "cli.md"))
    missing = [c for c in cli.leaf_commands() if f"### `shipyard {c}`" not in text]
    assert not missing, missing
    cfg = _read(os.path.join(root, "docs", "configuration.md"))
    for section in ("config.yaml", "credentials.yaml", "pool.yaml", "jobs.yaml", "fs.yaml", "federation.yaml", "monitor.yaml", "slurm.yaml"):
        assert f"## {section}" in cfg


def test_credentials_from_local_keyvault(tmp_path, monkeypatch):
    """`keyvault add` stores the credentials section; `--keyvault-credentials-secret-id` then replaces credentials.yaml."""
    monkeypatch.setenv("SHIPYARD_STATE_DIR", str(tmp_path / "state"))
    cfg = tmp_path / "cfg"
    cfg.mkdir()
    (cfg / "credentials.yaml").write_text("credentials:\n  storage:\n    acct: {account: local}\n")
    (cfg / "config.yaml").write_text("batch_shipyard: {storage_account_settings: acct}\nglobal_resources: {docker_images: [busybox]}\n")
    (cfg / "pool.yaml").write_text("pool_specification: {id: kv, vm_size: STANDARD_D2_V2, vm_count: {dedicated: 1, low_priority: 0}}\n")
    r = CliRunner()
    res = r.invoke(cli.cli, ["keyvault", "add", "mycreds", "--configdir", str(cfg), "--raw"], obj=cli.CliContext())
    assert res.exit_code == 0, res.output
    assert "mycreds" in r.invoke(cli.cli, ["keyvault", "list", "--raw"], obj=cli.CliContext()).output
    (cfg / "credentials.yaml").unlink()                               # from here on the vault is the only source of credentials
    res = r.invoke(cli.cli, ["pool", "list", "--configdir", str(cfg), "--raw", "--show-config"], obj=cli.CliContext())
    assert res.exit_code == 0 and "acct: {" not in res.output and "account: local" not in res.output   # no credentials anywhere now
    res = r.invoke(cli.cli, ["pool", "list", "--configdir", str(cfg), "--raw", "--keyvault-credentials-secret-id", "mycreds"], obj=cli.CliContext())
    assert res.exit_code == 0, res.output
    res = r.invoke(cli.cli, ["pool", "list", "--configdir", str(cfg), "--raw", "--keyvault-credentials-secret-id", "https://local.vault/secrets/mycreds",
                             "--show-config"], obj=cli.CliContext())
    assert res.exit_code == 0 and "account: local" in res.output       # the merged config really contains the vault's credentials
    res = r.invoke(cli.cli, ["pool", "list", "--configdir", str(cfg), "--keyvault-credentials-secret-id", "nope"], obj=cli.CliContext())
    assert res.exit_code == 1 and "not found" in res.output


def test_slurm_helper_node_side_verbs(tmp_path, capsys):
    """The reference helper's verbs and host options: --hostfile ("host partition" lines), --host, sakey, get-node-assignment,
    complete-node-assignment, check-provisioning-status."""
    import yaml
    from batch_shipyard_b200.backend.local import LocalBackend
    from batch_shipyard_b200.slurm import cluster as sl
    sd = str(tmp_path / "st")
    b = LocalBackend(state_dir=sd)
    conf = tmp_path / "slurm.yaml"
    conf.write_text(yaml.safe_dump({"slurm": {"cluster_id": "sc", "slurm_options": {"elastic_partitions": {}}}}))
    b.store.insert("slurmhost", "sc", "sc-p-x-0", {"name": "sc-p-x-0", "state": "up", "gpu_node": "gpu-3", "partition": "p", "pool": "x"})
    b.store.insert("slurmhost", "sc", "sc-p-x-1", {"name": "sc-p-x-1", "state": "suspended", "gpu_node": None, "partition": "p", "pool": "x"})
    base = ["--conf", str(conf), "--state-dir", sd]
    assert sl.main(["sakey"] + base) == 0 and "local" in capsys.readouterr().out
    assert sl.main(["get-node-assignment", "--host", "gpu-3"] + base) == 0                 # a node asks which Slurm host it is
    assert json.loads(capsys.readouterr().out)["host"] == "sc-p-x-0"
    assert sl.main(["get-node-assignment", "--host", "gpu-9"] + base) == 1
    capsys.readouterr()
    assert sl.main(["complete-node-assignment", "--host", "sc-p-x-0"] + base) == 0
    assert b.store.get("slurmhost", "sc", "sc-p-x-0")["assignment_complete"] is True
    capsys.readouterr()
    assert sl.main(["check-provisioning-status", "--host", "sc-p-x-0"] + base) == 0
    assert sl.main(["check-provisioning-status", "--host", "sc-p-x-1"] + base) == 1         # still suspended
    capsys.readouterr()
    hf = tmp_path / "hosts"
    hf.write_text("sc-p-x-1 p\n\nghost p\n")
    rc = sl.main(["resume-fail", "--hostfile", str(hf)] + base)
    out = json.loads(capsys.readouterr().out)
    assert rc in (0, 1) and isinstance(out, dict)


def test_no_command_leaks_a_traceback(tmp_path, monkeypatch):
    """Every leaf command, invoked with plausible arguments against a provisioned pool, either succeeds or ends with `ERROR: ...`
    and exit code 1 — domain errors (unknown federation, missing storage cluster, ...) never surface as Python tracebacks."""
    import click
    monkeypatch.setenv("SHIPYARD_STATE_DIR", str(tmp_path / "state"))
    monkeypatch.setenv("SHIPYARD_INLINE_AGENT", "1")
    monkeypatch.setenv("SHIPYARD_FAKE_GPUS", "8")
    monkeypatch.chdir(tmp_path)                                   # `cert create` & co. write relative to the working directory
    recipe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "recipes", "mpiBench-OpenMPI", "config")
    r = CliRunner()
    assert r.invoke(cli.cli, ["pool", "add", "--configdir", recipe, "--raw", "-y"], obj=cli.CliContext()).exit_code == 0
    leaves = []

    def walk(g, path):
        for n, c in g.commands.items():
            (walk(c, path + [n]) if isinstance(c, click.Group) else leaves.append((path + [n], c)))
    walk(cli.cli, [])
    skip = {("pool", "del"), ("pool", "add"), ("jobs", "add"), ("storage", "del"), ("storage", "clear"), ("misc", "tensorboard")}
    plausible = {"federation_id": "fed1", "storage_cluster_id": "sc1", "name": "sec1", "storage_account": "acct", "path": "cont/x"}
    leaked = []
    for path, c in leaves:
        if tuple(path) in skip:
            continue
        args = list(path) + [plausible.get(p.name, "x") for p in c.params if isinstance(p, click.Argument) and p.nargs != -1]
        res = r.invoke(cli.cli, args + ["--configdir", recipe, "--raw", "-y"], obj=cli.CliContext())
        if res.exception is not None and not isinstance(res.exception, SystemExit):
            leaked.append((" ".join(path), repr(res.exception)[:160]))
        elif res.exit_code not in (0, 1, 2):
            leaked.append((" ".join(path), f"exit code {res.exit_code}"))
    assert leaked == [], leaked
    assert len(leaves) == 105


def test_quickstart_guide_example_runs_as_written(tmp_path, monkeypatch):
    """docs/quickstart.md: the four YAML snippets are split into files exactly as the comments say, then pool add / jobs add --tail."""
    import re
    doc = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "docs", "quickstart.md")).read()
    block = re.search(r"```yaml\n(# credentials\.yaml.*?)```", doc, flags=re.S).group(1)
    cfg = tmp_path / "cfg"
    cfg.mkdir()
    parts = re.split(r"^# (\w+)\.yaml\n", block, flags=re.M)[1:]
    names = parts[0::2]
    assert names == ["credentials", "config", "pool", "jobs"]
    for name, body in zip(names, parts[1::2]):
        (cfg / f"{name}.yaml").write_text(body)
    monkeypatch.setenv("SHIPYARD_STATE_DIR", str(tmp_path / "state"))
    monkeypatch.setenv("SHIPYARD_INLINE_AGENT", "1")
    monkeypatch.setenv("SHIPYARD_FAKE_GPUS", "2")
    from batch_shipyard_b200.pool import topology
    topology.probe(refresh=True)                       # the probe result is cached per process: re-read it under the fake-GPU setting
    try:
        r = CliRunner()
        res = r.invoke(cli.cli, ["pool", "add", "--configdir", str(cfg), "-y"], obj=cli.CliContext())
        assert res.exit_code == 0, res.output
        res = r.invoke(cli.cli, ["jobs", "add", "--configdir", str(cfg), "--tail", "stdout.txt"], obj=cli.CliContext())
        assert res.exit_code == 0 and "hello from gpu-" in res.output and "on GPU" in res.output, res.output
    finally:
        monkeypatch.delenv("SHIPYARD_FAKE_GPUS")
        topology.probe(refresh=True)
