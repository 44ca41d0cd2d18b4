"""State store primitives, cascade pre-load, staging library, data ingress."""
import os
import threading
import time

import pytest

from batch_shipyard_b200.data import ingress
from batch_shipyard_b200.ops.stage import Stager
from batch_shipyard_b200.pool import cascade as C
from batch_shipyard_b200.state.store import EntityExists, EtagMismatch, Store, entity_names


def test_store_entities_cas_queues_leases(tmp_path):
    st = Store(str(tmp_path))
    st.insert("k", "p", "r", {"a": 1})
    with pytest.raises(EntityExists):
        st.insert("k", "p", "r", {"a": 2})
    e = st.get("k", "p", "r")
    st.update("k", "p", "r", {"a": 3}, etag=e["_etag"])
    with pytest.raises(EtagMismatch):
        st.update("k", "p", "r", {"a": 4}, etag=e["_etag"])
    assert st.merge("k", "p", "r", {"b": 1}) == {"a": 3, "b": 1}
    assert len(st.query("k", "p")) == 1 and st.delete("k", "p") == 1
    mid = st.put_message("q", {"x": 1}); st.put_message("q", {"x": 2})
    m = st.get_messages("q", 1, visibility_timeout=0.2)[0]
    assert m["body"] == {"x": 1} and m["dequeue_count"] == 1 and m["id"] == mid
    assert [x["body"]["x"] for x in st.get_messages("q", 5, 0.2)] == [2]       # first one is invisible
    time.sleep(0.25)
    again = st.get_messages("q", 5, 1)
    assert [x["dequeue_count"] for x in again] == [2, 2]
    assert not st.delete_message("q", mid, pop_receipt="stale") and st.delete_message("q", mid, again[0]["pop_receipt"])
    assert st.acquire_lease("L", "a", 0.2) and not st.acquire_lease("L", "b", 0.2) and st.renew_lease("L", "a", 0.2)
    time.sleep(0.25)
    assert st.acquire_lease("L", "b", 1) and st.lease_holder("L") == "b" and not st.renew_lease("L", "a", 1)
    st.put_blob("c", "d/x.bin", b"123")
    assert st.get_blob("c", "d/x.bin") == b"123" and st.list_blobs("c") == ["d/x.bin"]
    with pytest.raises(ValueError):
        st.blob_path("c", "../../etc/passwd")
    assert entity_names("shipyard", "acct", "Pool1")["blob_globalresources"] == "shipyardgr-acct-pool1"
    with pytest.raises(ValueError):
        entity_names("x" * 60, "acct", "pool")


def test_store_concurrent_mutate(tmp_path):
    st = Store(str(tmp_path))
    st.insert("c", "p", "r", {"n": 0})

    def bump():
        s2 = Store(str(tmp_path))
        for _ in range(50):
            s2.mutate("c", "p", "r", lambda d: d.__setitem__("n", d["n"] + 1))
    ts = [threading.Thread(target=bump) for _ in range(4)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert st.get("c", "p", "r")["n"] == 200


def test_cascade_bounded_concurrency_retry_and_events(tmp_path):
    st = Store(str(tmp_path))
    res = [f"docker:img{i}" for i in range(6)]
    C.Cascade.populate(st, "p", res)
    live, peak, attempts, lock = [0], [0], {}, threading.Lock()

    def puller(r):
        with lock:
            attempts[r] = attempts.get(r, 0) + 1
            n = attempts[r]
            live[0] += 1; peak[0] = max(peak[0], live[0])
        time.sleep(0.05)
        with lock:
            live[0] -= 1
        if r == "docker:img2" and n < 3:
            raise RuntimeError("toomanyrequests: rate limited")            # transient -> retried with backoff
        if r == "docker:img4":
            raise RuntimeError("manifest unknown")                         # permanent
        return 1000
    cas = C.Cascade(st, "p", concurrency=2, puller=puller, sleep=lambda s: None)
    assert cas.run() is False
    assert peak[0] <= 2 and attempts["docker:img2"] == 3 and attempts["docker:img4"] == 1
    states = {e["resource"]: e["state"] for e in cas.resources()}
    assert states["docker:img4"] == "failed" and sum(1 for s in states.values() if s == "loaded") == 5
    ev = [e["event"] for e in st.events("p")]
    assert ev.count("pull-start") == 6 and ev.count("pull-end") == 5 and "gr-done" not in ev
    assert not C.wait_for_images(st, "p", timeout=0.3)
    # fallback registry rescues a permanent failure
    C.Cascade.populate(st, "q", ["docker:only-in-mirror"])
    seen = []

    def p2(r):
        seen.append(r)
        if r.startswith("docker:mirror.io/"):
            return 5
        raise RuntimeError("manifest unknown")
    assert C.Cascade(st, "q", concurrency=1, puller=p2, fallback_registry="mirror.io").run()
    assert seen == ["docker:only-in-mirror", "docker:mirror.io/only-in-mirror"] and C.wait_for_images(st, "q", 1)
    d = list(zip(range(12), C.backoff_delays()))
    assert all(0 <= x <= 300 for _, x in d)


def test_cascade_stages_local_artefact(tmp_path, monkeypatch):
    st = Store(str(tmp_path / "s"))
    img = tmp_path / "images"
    img.mkdir()
    (img / "docker-busybox.tar").write_bytes(os.urandom(3_000_000))
    monkeypatch.setenv("SHIPYARD_IMAGE_DIR", str(img))
    C.Cascade.populate(st, "p", ["docker:busybox"])
    assert C.Cascade(st, "p", concurrency=2).run()
    e = st.query("globalresource", "p")[0]
    assert e["state"] == "loaded" and e["size"] == 3_000_000
    monkeypatch.setenv("SHIPYARD_STRICT_IMAGES", "1")
    C.Cascade.populate(st, "p2", ["docker:absent"])
    assert not C.Cascade(st, "p2", concurrency=1).run()


def test_stager_host_mode(tmp_path):
    s = Stager(None, arena_bytes=4 << 20, concurrency=3)
    blobs = {}
    for i in range(5):
        p = tmp_path / f"f{i}.bin"
        blobs[i] = os.urandom(700_001 * (i + 1))
        p.write_bytes(blobs[i])
    tickets = {i: s.submit_file(str(tmp_path / f"f{i}.bin")) for i in range(5)}
    part = s.submit_file(str(tmp_path / "f4.bin"), offset=100, nbytes=1000)
    for i, t in tickets.items():
        assert s.wait(t, 30) and s.read_host(t) == blobs[i]
    assert s.wait(part, 30) and s.read_host(part) == blobs[4][100:1100]
    st = s.stats()
    assert st["bytes_staged"] == sum(len(b) for b in blobs.values()) + 1000
    with pytest.raises(Exception):
        s.submit_file(str(tmp_path / "missing.bin"))
    s.close()


def test_ingress_bin_packing_split_and_filters(tmp_path):
    src = tmp_path / "src"
    (src / "sub").mkdir(parents=True)
    sizes = {"a.dat": 900, "b.dat": 500, "sub/c.dat": 400, "d.bak": 100, "e.dat": 300}
    for n, sz in sizes.items():
        (src / n).write_bytes(os.urandom(sz))
    ents = ingress.walk_source(str(src), include=["*.dat"], exclude=["e.*"])
    assert sorted(e.rel for e in ents) == ["a.dat", "b.dat", "sub/c.dat"]
    buckets = ingress.bin_pack(ents, ["n0", "n1"])
    assert sorted(b.bytes for b in buckets) == [900, 900]                       # 900 | 500+400
    big = tmp_path / "big.bin"
    data = os.urandom((3 << 20) + 12345)
    big.write_bytes(data)
    parts = ingress.split_entries(ingress.walk_source(str(big)), 1)
    assert len(parts) == 4 and sum(p.size for p in parts) == len(data)
    stats = ingress.transfer(ingress.bin_pack(parts, ["n0", "n1", "n2"]), str(tmp_path / "dst"), 2)
    assert (tmp_path / "dst" / "big.bin").read_bytes() == data and stats["bytes"] == len(data) and stats["mbit_per_s"] > 0


def test_perfgraph_coalesces_intervals(tmp_path):
    from batch_shipyard_b200.misc import perfgraph
    from batch_shipyard_b200.state.store import Store
    st = Store(str(tmp_path / "state"))
    t = 1000.0
    st.record_event("nodeprep", "start", pool="p", node="n0", ts=t)
    st.record_event("cascade", "start", pool="p", node="n0", ts=t + 1)
    st.record_event("cascade", "pull-start", pool="p", node="n0", message="img:a", ts=t + 1.5)
    st.record_event("cascade", "pull-end", pool="p", node="n0", message="img:a,size=10,seconds=0.5", ts=t + 2)
    st.record_event("cascade", "gr-done", pool="p", node="n0", message="nglobalresources=1", ts=t + 2.5)
    st.record_event("nodeprep", "end", pool="p", node="n0", ts=t + 3)
    tl = perfgraph.coalesce(st.events("p"))
    iv = {name: (round(a, 3), round(b, 3)) for name, a, b in tl["n0"]["intervals"]}
    assert iv["nodeprep:run"] == (0.0, 3.0)
    assert iv["cascade:run"] == (1.0, 2.5)
    assert iv["cascade:pull[img:a]"] == (1.5, 2.0)
    text = perfgraph.render_text(tl)
    assert "cascade:pull[img:a]" in text and "#" in text
    dat, gp = perfgraph.render_gnuplot(tl, str(tmp_path / "tl"))
    assert open(dat).read().count("\n") == 3 and "boxxyerrorbars" in open(gp).read()
    assert perfgraph.main(["--state-dir", str(tmp_path / "state"), "--pool", "p", "--format", "json"]) == 0


def test_mover_refuses_remote_paths_that_leave_the_storage_root(tmp_path, capsys):
    from batch_shipyard_b200.data import mover
    sd = str(tmp_path / "state")
    root = mover.storage_root(sd, "acct")
    os.makedirs(os.path.join(root, "cont"), exist_ok=True)
    with open(os.path.join(root, "cont", "a.txt"), "w") as f:
        f.write("x")
    secret = tmp_path / "secret.txt"
    secret.write_text("s")
    assert mover.remote_path(root, "cont/sub") == os.path.join(root, "cont", "sub")
    assert mover.main(["ingress", "--state-dir", sd, "--link", "acct", "--remote", "cont", "--local", str(tmp_path / "in")]) == 0
    assert (tmp_path / "in" / "a.txt").read_text() == "x"
    rel = os.path.relpath(str(tmp_path), root)                                    # ../../.. up to the directory holding secret.txt
    assert mover.main(["ingress", "--state-dir", sd, "--link", "acct", "--remote", rel, "--local", str(tmp_path / "leak")]) == 1
    assert not (tmp_path / "leak").exists() and "leaves the storage account" in capsys.readouterr().err
    os.environ["SHIPYARD_TASK_RESULT"] = "success"
    try:
        assert mover.main(["egress", "--state-dir", sd, "--link", "acct", "--remote", "../outside", "--local", str(tmp_path / "in")]) == 1
    finally:
        os.environ.pop("SHIPYARD_TASK_RESULT", None)
    assert not os.path.exists(os.path.join(os.path.dirname(root), "outside"))


def test_copy_tree_uses_the_native_parallel_copier_and_matches_shutil(tmp_path, monkeypatch):
    """Files of 1 MB and more go through sy_stage_submit_copy on worker threads (content, mode and mtime as shutil.copy2), small ones inline;
    include / exclude filters, the collected path list and the failure path (unwritable destination) behave the same with and without it."""
    import os
    from batch_shipyard_b200.data import mover
    from batch_shipyard_b200.ops import stage as stage_mod
    src = tmp_path / "src"; (src / "sub").mkdir(parents=True)
    blobs = {"big0.bin": os.urandom((3 << 20) + 17), "sub/big1.bin": os.urandom(1 << 20), "small.txt": b"hello", "sub/skip.tmp": os.urandom(2 << 20)}
    for rel, data in blobs.items():
        p = src / rel; p.write_bytes(data); os.chmod(p, 0o640); os.utime(p, (1_600_000_000, 1_600_000_500))
    calls = []
    real = stage_mod.Stager.submit_copy
    monkeypatch.setattr(stage_mod.Stager, "submit_copy", lambda self, a, b, *k: calls.append(a) or real(self, a, b, *k))
    for native in ("1", "0"):
        monkeypatch.setenv("SHIPYARD_NATIVE_COPY", native)
        calls.clear()
        dst = tmp_path / f"dst{native}"
        got = []
        n, nb = mover.copy_tree(str(src), str(dst), exclude=["*.tmp"], collect=got)
        assert (n, nb) == (3, sum(len(v) for k, v in blobs.items() if not k.endswith(".tmp"))) and len(got) == 3
        assert sorted(os.path.basename(c) for c in calls) == (["big0.bin", "big1.bin"] if native == "1" else [])
        for rel, data in blobs.items():
            q = dst / rel
            if rel.endswith(".tmp"):
                assert not q.exists()
                continue
            st = os.stat(q)
            assert q.read_bytes() == data and (st.st_mode & 0o777) == 0o640 and int(st.st_mtime) == 1_600_000_500, rel
    # a failing native copy surfaces as an error, not as a silently missing file
    monkeypatch.setenv("SHIPYARD_NATIVE_COPY", "1")
    ro = tmp_path / "ro"; ro.mkdir(); (ro / "big0.bin").mkdir()          # destination path is a directory: open(O_WRONLY) fails
    import pytest as _pytest
    with _pytest.raises(OSError):
        mover.copy_tree(str(src / "big0.bin"), str(ro))
