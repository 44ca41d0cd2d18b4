"""known bugs / quirks of the reference, kept as regression tests that this build does NOT replicate them.

Q2 and Q11 are covered in tests/test_config.py and tests/test_mpi_autoscale.py; Q5 (executor.submit(f(...))), Q7 (print(creds)) and
Q10 (vestigial torrent containers) have no counterpart to test against (there is no such code here)."""
import copy
import subprocess

from _helpers import make, up

from batch_shipyard_b200.config import settings as S
from batch_shipyard_b200.jobs import builder, mpi as M
from batch_shipyard_b200.utils import util


def test_q1_idle_count_is_the_idle_count(tmp_path):
    """reference: `pool nodes count` prints the *creating* count on the `idle:` line (convoy/batch.py:2902,2920)."""
    cfg, b = make(tmp_path)
    up(cfg, b)
    counts = b.node_counts("testpool")
    states = [n["state"] for n in b.list_nodes("testpool")]
    assert counts["dedicated"]["idle"] == states.count("idle") == 2
    assert counts["dedicated"]["creating"] == 0 and counts["dedicated"]["total"] == 2


def test_q3_env_dump_without_exclusions_redirects(tmp_path):
    """reference: with nothing to exclude it emits `env | <file>` (pipes into the file name, convoy/batch.py:4348)."""
    line = builder.env_dump_command(exclude=(), env_file=str(tmp_path / "envlist"))
    assert "|" not in line and ">" in line
    subprocess.run(["/bin/bash", "-c", line], check=True, env={"FOO": "bar", "PATH": "/usr/bin:/bin"})
    assert "FOO=bar" in (tmp_path / "envlist").read_text()
    excl = builder.env_dump_command(exclude=("SECRET",), env_file=str(tmp_path / "e2"))
    subprocess.run(["/bin/bash", "-c", excl], check=True, env={"SECRET": "x", "KEEP": "y", "PATH": "/usr/bin:/bin"})
    txt = (tmp_path / "e2").read_text()
    assert "KEEP=y" in txt and "SECRET" not in txt


def _task(cfg, jobspec_extra=None, task_extra=None):
    pool = S.pool_settings(cfg)
    jobspec = copy.deepcopy(cfg["job_specifications"][0])
    jobspec.update(jobspec_extra or {})
    task = dict(jobspec["tasks"][0]); task.update(task_extra or {})
    return builder.build_task(cfg, pool, jobspec, task, "t0", dry_run=True)


def test_q4_job_level_infiniband_is_honoured(tmp_path):
    """reference: the job-level `infiniband` value is read and then dropped (convoy/settings.py:4254-4255)."""
    cfg, _ = make(tmp_path)
    assert _task(cfg).infiniband is False
    assert _task(cfg, jobspec_extra={"infiniband": True}).infiniband is True
    assert _task(cfg, jobspec_extra={"infiniband": True}, task_extra={"infiniband": False}).infiniband is False   # task overrides job


def test_q6_wrapped_commands_use_valid_shell_options():
    """reference: `set e+` instead of `set +e` in wrap_commands (convoy/util.py:364)."""
    w = util.wrap_commands_in_shell(["false", "echo not-reached"])
    assert "set e+" not in w and "set -e" in w
    p = subprocess.run(w, shell=True, stdout=subprocess.PIPE, text=True)
    assert p.returncode != 0 and "not-reached" not in p.stdout               # fail-fast really is in effect
    ok = subprocess.run(util.wrap_commands_in_shell(["echo a", "echo b"]), shell=True, stdout=subprocess.PIPE, text=True)
    assert ok.returncode == 0 and ok.stdout.split() == ["a", "b"]


def test_q8_mpi_runtime_names_follow_schema_not_docs():
    """reference docs say `intelmpi_ofa` and omit `mvapich`; schema and code use `intelmpi-ofa` and accept `mvapich`."""
    assert "intelmpi-ofa" in M.RUNTIMES and "mvapich" in M.RUNTIMES and "intelmpi_ofa" not in M.RUNTIMES


def test_q9_remove_container_after_exit_inherits_job_then_defaults_true(tmp_path):
    """reference docs claim the task-level default is false; the code inherits the job value, then defaults to true."""
    cfg, _ = make(tmp_path)
    assert "--rm" in _task(cfg).container_command
    assert "--rm" not in _task(cfg, jobspec_extra={"remove_container_after_exit": False}).container_command
    assert "--rm" in _task(cfg, jobspec_extra={"remove_container_after_exit": False},
                           task_extra={"remove_container_after_exit": True}).container_command


def test_q12_federated_job_schedule_reports_tasks_per_recurrence(tmp_path):
    """reference: caller passes kind 'job_schedule', the builder tests 'jobschedule', so tasks_per_recurrence is never set."""
    from batch_shipyard_b200.jobs import submit
    cfg, b = make(tmp_path, job={"auto_complete": True, "recurrence": {"schedule": {"recurrence_interval": "00:10:00"}}},
                  tasks=[{"docker_image": "busybox", "command": "echo 1"}, {"docker_image": "busybox", "command": "echo 2"}])
    up(cfg, b)
    out = submit.add_jobs(b, cfg)
    summary = out["job1"] if "job1" in out else list(out.values())[0]
    assert summary.get("kind") == "job_schedule" and summary.get("tasks_per_recurrence") == 2


def test_q13_slurm_shared_volume_mount_path_spellings():
    """reference: schema and code read `host_mount_path`, its own template and docs write `mount_path` (the template fails the
    reference's schema).  Both spellings validate here and resolve to the same setting; neither is an error that names the key."""
    import pytest
    from batch_shipyard_b200.config.schema import ConfigType, ValidationError, validate
    base = {"slurm": {"cluster_id": "c", "storage_account_settings": "acct", "location": "x", "resource_group": "rg",
                      "shared_data_volumes": {"nfs": {"store_slurmctld_state": True}},
                      "slurm_options": {"elastic_partitions": {"p": {"batch_pools": {"pool": {"max_compute_nodes": 2}}}}}}}
    for key in ("host_mount_path", "mount_path"):
        cfg = copy.deepcopy(base)
        cfg["slurm"]["shared_data_volumes"]["nfs"][key] = "/shared"
        try:
            validate(ConfigType.Slurm, cfg)
        except ValidationError as e:                      # other required keys of the full schema are not the point here
            assert not any("mount_path" in m for m in e.errors), e.errors
        assert S.slurm_options(cfg)["shared_data_volumes"]["nfs"] == {"host_mount_path": "/shared", "store_slurmctld_state": True}
    with pytest.raises(ValueError, match="host_mount_path"):
        S.slurm_options(base)
    bad = copy.deepcopy(base); bad["slurm"]["shared_data_volumes"]["nfs"]["mount_path"] = "/home/"
    with pytest.raises(ValueError, match="/home"):
        S.slurm_options(bad)
