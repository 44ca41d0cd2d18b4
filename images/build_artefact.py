#!/usr/bin/env python3
"""Build a local image artefact: the one-box counterpart of `docker build` + `docker push` for the images the recipes name.

The reference builds its images with Dockerfiles under images/ and cascade/ and the pool pulls them from a registry
(/root/reference/images/docker/linux/cli/Dockerfile, /root/reference/cascade/cascade.py:500-571).  This box has no container runtime and
no registry, so an "image" is an artefact in the pool's image store (``<state>/images`` or $SHIPYARD_IMAGE_DIR): a tar archive of the
shipyard runtime (package + native libraries + the named recipe bodies) with a manifest.  ``shipyard pool add`` then really pre-loads it:
cascade finds the artefact for every ``global_resources.docker_images`` entry, streams it through the native stager with bounded
concurrency (lease slots, retries, perf events) and ``jobs add`` accepts tasks on that image.

    python images/build_artefact.py --image shipyard/pytorch:b200 --recipe PyTorch-GPU [--state-dir DIR | --image-dir DIR]
"""
import argparse
import hashlib
import io
import json
import os
import sys
import tarfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _files(recipes):
    keep = []
    for base in ["batch_shipyard_b200", "shipyard"] + [os.path.join("recipes", r) for r in recipes]:
        p = os.path.join(ROOT, base)
        if os.path.isfile(p):
            keep.append(base)
            continue
        for d, dirs, fs in os.walk(p):
            dirs[:] = [x for x in dirs if x != "__pycache__"]
            for f in fs:
                if not f.endswith((".pyc", ".o")):
                    keep.append(os.path.relpath(os.path.join(d, f), ROOT))
    return sorted(keep)


def build(image: str, recipes, image_dir: str, kind: str = "docker") -> dict:
    from batch_shipyard_b200.pool.cascade import artefact_name
    os.makedirs(image_dir, exist_ok=True)
    out = os.path.join(image_dir, artefact_name(f"{kind}:{image}") + ".tar")
    files = _files(recipes)
    manifest = {"image": image, "kind": kind, "built": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "recipes": list(recipes), "files": {}}
    tmp = out + ".tmp"
    with tarfile.open(tmp, "w") as tar:
        for rel in files:
            p = os.path.join(ROOT, rel)
            with open(p, "rb") as f:
                manifest["files"][rel] = hashlib.sha256(f.read()).hexdigest()
            tar.add(p, arcname=os.path.join("opt/shipyard-b200", rel), recursive=False)
        blob = json.dumps(manifest, indent=1).encode()
        info = tarfile.TarInfo("opt/shipyard-b200/IMAGE_MANIFEST.json"); info.size = len(blob); info.mtime = int(time.time())
        tar.addfile(info, io.BytesIO(blob))
    os.replace(tmp, out)
    return {"artefact": out, "bytes": os.path.getsize(out), "files": len(files), "image": image}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--image", required=True, help="image name as written in global_resources.docker_images, e.g. shipyard/pytorch:b200")
    ap.add_argument("--recipe", action="append", default=[], help="recipe directory to include (repeatable)")
    ap.add_argument("--kind", default="docker", choices=["docker", "singularity"])
    ap.add_argument("--state-dir", default=os.environ.get("SHIPYARD_STATE_DIR"))
    ap.add_argument("--image-dir", default=os.environ.get("SHIPYARD_IMAGE_DIR"))
    a = ap.parse_args(argv)
    if not a.image_dir:
        from batch_shipyard_b200.state.store import default_state_dir
        a.image_dir = os.path.join(a.state_dir or default_state_dir(), "images")
    print(json.dumps(build(a.image, a.recipe, a.image_dir, a.kind)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
