"""Sanitizer runs of the host-side native code: the multi-threaded staging library under ThreadSanitizer,
the task runner under Address + UndefinedBehaviour sanitizers driving a real multi-instance task."""
import importlib.util
import os
import subprocess

import pytest

from _helpers import make, read, run, up

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def san():
    spec = importlib.util.spec_from_file_location("shipyard_native_build", os.path.join(ROOT, "native", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        return mod.build_sanitizers()
    except RuntimeError as e:                     # toolchain without sanitizer runtimes
        pytest.skip(f"sanitizer build unavailable: {e}")


def test_stage_library_is_race_free_under_tsan(san, tmp_path):
    f = tmp_path / "blob.bin"
    f.write_bytes(os.urandom(700_000))
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66")
    p = subprocess.run([san["stage_tsan"], str(f)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0 and "ThreadSanitizer" not in p.stdout, p.stdout[-3000:]
    assert "0 failures" in p.stdout


def test_task_runner_clean_under_asan_ubsan(san, tmp_path, monkeypatch):
    monkeypatch.setenv("SHIPYARD_TASKRUN_BIN", san["taskrun_asan"])
    monkeypatch.setenv("ASAN_OPTIONS", "detect_leaks=0:halt_on_error=1:exitcode=67")
    monkeypatch.setenv("UBSAN_OPTIONS", "halt_on_error=1:print_stacktrace=1")
    mi = {"num_instances": 2, "mpi": {"runtime": "openmpi", "processes_per_node": 1}}
    tasks = [{"id": "mi", "docker_image": "busybox", "multi_instance": mi, "environment_variables": {"FOO": "bar baz"},
              "command": "/bin/sh -c 'echo rank=$RANK/$WORLD_SIZE foo=$FOO'"},
             {"id": "plain", "docker_image": "busybox", "command": "/bin/sh -c 'echo plain; exit 3'", "max_task_retries": 1}]
    cfg, b = make(tmp_path, tasks=tasks)
    up(cfg, b)
    run(cfg, b)
    out = read(b, "job1", "mi")
    assert "rank=0/2 foo=bar baz" in out
    for name in ("stderr.txt",):
        err = read(b, "job1", "mi", name)
        assert "AddressSanitizer" not in err and "runtime error" not in err, err[-2000:]
    t = b.get_task("job1", "plain")
    assert t["exit_code"] == 3                                        # the runner's own exit-code contract survives the sanitizer build
    err = read(b, "job1", "plain", "stderr.txt")
    assert "AddressSanitizer" not in err and "runtime error" not in err, err[-2000:]
