"""shipyard-diskbench (native/bench/diskbench.cpp): the DiskSpd-style storage benchmark behind recipes/DiskSpd
(reference: /root/reference/recipes/DiskSpd-Windows/config/jobs.yaml, command `-c8192k -d1 testfile.dat`)."""
import json
import os
import subprocess

import pytest

from batch_shipyard_b200._build import ensure_built, native_dir


@pytest.fixture(scope="module")
def exe():
    ensure_built(["diskbench"])
    p = os.path.join(native_dir(), "shipyard-diskbench")
    assert os.path.exists(p)
    return p


def _run(exe, args, cwd):
    return subprocess.run([exe, *args], cwd=cwd, capture_output=True, text=True, timeout=120)


def test_reference_command_line_creates_reads_and_cleans_up(exe, tmp_path):
    r = _run(exe, ["-c8192k", "-d0.3", "testfile.dat"], tmp_path)
    assert r.returncode == 0, r.stderr
    assert "testfile.dat (8388608 bytes)" in r.stdout and "block 65536" in r.stdout and "errors 0" in r.stdout
    total = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("total")][0]
    assert int(total[1]) > 0 and int(total[1]) == int(total[2]) * 65536              # bytes == I/Os * block
    assert not (tmp_path / "testfile.dat").exists()                                 # a file the run created is removed (-k keeps it)


def test_json_random_mixed_threads_and_keep(exe, tmp_path):
    r = _run(exe, ["-c4m", "-d0.3", "-W0.1", "-b4k", "-t4", "-w30", "-r", "-j", "-k", "f.dat"], tmp_path)
    assert r.returncode == 0, r.stderr
    j = json.loads(r.stdout)
    assert j["file_bytes"] == 4 << 20 and j["block_bytes"] == 4096 and j["threads"] == 4 and j["pattern"] == "random" and j["write_pct"] == 30
    assert j["ios"] == j["read_ios"] + j["write_ios"] and j["bytes"] == j["ios"] * 4096 and j["errors"] == 0
    assert 0.15 < j["write_ios"] / j["ios"] < 0.45                                  # ~30 % writes
    lat = j["lat_us"]
    assert 0 < lat["p50"] <= lat["p95"] <= lat["p99"] and lat["max"] > 0
    assert abs(j["mb_per_s"] - j["bytes"] / j["seconds"] / 1e6) < 0.02 * j["mb_per_s"] + 0.01
    assert (tmp_path / "f.dat").stat().st_size == 4 << 20                           # -k
    before = (tmp_path / "f.dat").read_bytes()
    r = _run(exe, ["-d0.2", "-b64k", "-j", "f.dat"], tmp_path)                      # existing file, read-only: content untouched, file kept
    assert r.returncode == 0 and json.loads(r.stdout)["write_ios"] == 0
    assert (tmp_path / "f.dat").read_bytes() == before
    r = _run(exe, ["-d0.2", "-b64k", "-w", "-j", "f.dat"], tmp_path)                # bare -w = 100 % writes
    assert r.returncode == 0 and json.loads(r.stdout)["read_ios"] == 0


def test_bad_usage_is_reported_not_crashed(exe, tmp_path):
    assert _run(exe, [], tmp_path).returncode == 2
    assert _run(exe, ["-z", "f"], tmp_path).returncode == 2
    assert _run(exe, ["-bnope", "f"], tmp_path).returncode == 2
    assert _run(exe, ["-S", "-b1000", "-c1m", "f"], tmp_path).returncode == 2      # O_DIRECT needs 4 KiB multiples
    r = _run(exe, ["missing.dat"], tmp_path)
    assert r.returncode == 1 and "use -c<size>" in r.stderr
    r = _run(exe, ["-c64k", "-b64k", "-t4", "-d0.1", "small.dat"], tmp_path)        # fewer blocks than threads
    assert r.returncode == 1 and "at least one per thread" in r.stderr and not (tmp_path / "small.dat").exists()


def test_direct_io_when_the_file_system_supports_it(exe, tmp_path):
    r = _run(exe, ["-c1m", "-b4k", "-S", "-d0.2", "-j", "d.dat"], tmp_path)
    if r.returncode != 0:
        assert "O_DIRECT" in r.stderr or "Invalid argument" in r.stderr               # tmpfs / overlay without O_DIRECT: a clear message
        return
    assert json.loads(r.stdout)["direct"] is True
