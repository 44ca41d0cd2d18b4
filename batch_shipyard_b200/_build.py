"""On-demand native builds (thin wrapper over native/build.py)."""
from __future__ import annotations

import importlib.util
import os

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _builder():
    path = os.path.join(_ROOT, "native", "build.py")
    spec = importlib.util.spec_from_file_location("shipyard_native_build", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ensure_built(names=None, force: bool = False, quiet: bool = True) -> None:
    _builder().build_all(names, force=force, quiet=quiet)


def native_path(name: str) -> str:
    return _builder().artifact(name)


def native_dir() -> str:
    return os.path.join(_ROOT, "batch_shipyard_b200", "_native")
