"""images/build_artefact.py: the local counterpart of building + pushing the recipes' container images; cascade pre-loads the result."""
import json
import os
import subprocess
import sys
import tarfile

from _helpers import make, up

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_artefact_and_cascade_preloads_it(tmp_path, monkeypatch):
    cfg, b = make(tmp_path, extra={"global_resources": {"docker_images": ["busybox", "shipyard/pytorch:b200"]}})
    monkeypatch.setenv("SHIPYARD_STRICT_IMAGES", "0")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "images", "build_artefact.py"), "--image", "shipyard/pytorch:b200",
                        "--recipe", "PyTorch-GPU", "--state-dir", b.root], stdout=subprocess.PIPE, text=True, check=True)
    res = json.loads(p.stdout)
    assert res["bytes"] > 100_000 and res["files"] > 50 and os.path.exists(res["artefact"])
    with tarfile.open(res["artefact"]) as tar:
        names = tar.getnames()
        man = json.load(tar.extractfile("opt/shipyard-b200/IMAGE_MANIFEST.json"))
    assert "opt/shipyard-b200/recipes/PyTorch-GPU/ddp_resnet50_stock.py" in names and "opt/shipyard-b200/batch_shipyard_b200/cli.py" in names
    assert man["image"] == "shipyard/pytorch:b200" and len(man["files"]) == res["files"]
    up(cfg, b)                                             # pool add: cascade finds the artefact and reads it through the native stager
    rows = {r["resource"]: r for r in b.store.query("globalresource", "testpool")}
    assert rows["docker:shipyard/pytorch:b200"]["state"] == "loaded" and rows["docker:shipyard/pytorch:b200"]["size"] == res["bytes"]
    ev = [(e["source"], e["event"], e.get("message") or "") for e in b.store.events("testpool")]
    assert any(s == "cascade" and e == "stage" and f"bytes={res['bytes']}" in m for s, e, m in ev)      # the stager moved every byte
