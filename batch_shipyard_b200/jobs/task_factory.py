"""Task factories: expand one task template into many concrete tasks.

Kinds and semantics follow the reference's documented behaviour
(/root/reference/docs/35-batch-shipyard-task-factory-merge-task.md:52-540,
implementation /root/reference/convoy/task_factory.py:305-464):
``parametric_sweep`` {product, product_iterables, combinations, permutations,
zip}, ``random`` {integer | 9 real-valued distributions}, ``repeat``, ``file``
(enumerate a storage location -> one task per file, with keyword formatters and
the file staged into the task), and ``custom`` (import a module and call its
``generate(*args, **kwargs)`` generator).  Generated values are applied to the
``command`` with ``str.format`` (``{}`` / ``{0}`` / keywords).

On the local backend a "storage account" is a directory
(``credentials.storage.<link>.local_path``), so ``file`` walks the filesystem.
"""
from __future__ import annotations

import copy
import fnmatch
import importlib
import itertools
import os
import random as _random
from typing import Iterator


def _tmpl(task: dict) -> dict:
    t = copy.deepcopy(task)
    t.pop("task_factory", None)
    return t


def _apply(task: dict, args: tuple = (), kwargs: dict | None = None) -> dict:
    t = _tmpl(task)
    cmd = t.get("command")
    if cmd is not None:
        t["command"] = str(cmd).format(*args, **(kwargs or {}))
    return t


def _ranges(specs: list[dict]) -> list[range]:
    out = []
    for s in specs:
        if int(s["step"]) == 0:
            raise ValueError("parametric_sweep.product step may not be 0")
        out.append(range(int(s["start"]), int(s["stop"]), int(s["step"])))
    return out


def _parametric(task: dict, sweep: dict) -> Iterator[dict]:
    if sweep.get("product") is not None:
        for combo in itertools.product(*_ranges(sweep["product"])):
            yield _apply(task, combo)
    elif sweep.get("product_iterables") is not None:
        for combo in itertools.product(*sweep["product_iterables"]):
            yield _apply(task, combo)
    elif sweep.get("combinations") is not None:
        c = sweep["combinations"]
        fn = itertools.combinations_with_replacement if c.get("replacement", False) else itertools.combinations
        for combo in fn(c["iterable"], int(c["length"])):
            yield _apply(task, combo)
    elif sweep.get("permutations") is not None:
        p = sweep["permutations"]
        for combo in itertools.permutations(p["iterable"], int(p["length"])):
            yield _apply(task, combo)
    elif sweep.get("zip") is not None:
        for combo in zip(*sweep["zip"]):
            yield _apply(task, combo)
    else:
        raise ValueError("parametric_sweep needs one of product, product_iterables, combinations, "
                         "permutations, zip")


_DISTS = {
    "uniform": lambda r, p: r.uniform(p["a"], p["b"]),
    "triangular": lambda r, p: r.triangular(p["low"], p["high"], p.get("mode")),
    "beta": lambda r, p: r.betavariate(p["alpha"], p["beta"]),
    "exponential": lambda r, p: r.expovariate(p["lambda"]),
    "gamma": lambda r, p: r.gammavariate(p["alpha"], p["beta"]),
    "gauss": lambda r, p: r.gauss(p["mu"], p["sigma"]),
    "lognormal": lambda r, p: r.lognormvariate(p["mu"], p["sigma"]),
    "pareto": lambda r, p: r.paretovariate(p["alpha"]),
    "weibull": lambda r, p: r.weibullvariate(p["alpha"], p["beta"]),
}


def _random_factory(task: dict, spec: dict) -> Iterator[dict]:
    rng = _random.Random(spec.get("seed")) if spec.get("seed") is not None else _random.Random()
    n = int(spec["generate"])
    integer, dist = spec.get("integer"), spec.get("distribution")
    if integer is not None:
        for _ in range(n):
            yield _apply(task, (rng.randrange(int(integer["start"]), int(integer["stop"]), int(integer["step"])),))
        return
    if dist:
        for name, fn in _DISTS.items():
            if dist.get(name) is not None:
                for _ in range(n):
                    yield _apply(task, (fn(rng, dist[name]),))
                return
    raise ValueError("random task factory needs 'integer' or one 'distribution'")


def enumerate_storage_files(config: dict, spec: dict) -> Iterator[dict]:
    """Yield keyword dicts for each file under the (local) storage location."""
    from ..config import settings
    az = spec["azure_storage"]
    link = az["storage_account_settings"]
    root = settings.credentials_storage_local_path(config, link)
    if root is None:
        raise ValueError(f"task_factory.file: credentials.storage.{link}.local_path is not configured "
                         "(a local directory stands in for the storage account)")
    remote = str(az["remote_path"]).strip("/")
    container, _, prefix = remote.partition("/")
    base = os.path.join(root, container)
    inc, exc = az.get("include") or [], az.get("exclude") or []
    if not os.path.isdir(base):
        return
    found = []
    for d, _, files in os.walk(base):
        for fn in files:
            rel = os.path.relpath(os.path.join(d, fn), base).replace(os.sep, "/")
            if prefix and not rel.startswith(prefix.rstrip("/") + "/") and rel != prefix:
                continue
            if inc and not any(fnmatch.fnmatch(rel, pat) for pat in inc):
                continue
            if any(fnmatch.fnmatch(rel, pat) for pat in exc):
                continue
            found.append(rel)
    for rel in sorted(found):
        name = rel.rsplit("/", 1)[-1]
        yield {"url": "file://" + os.path.join(base, rel), "file_path_with_container": f"{container}/{rel}",
               "file_path": rel, "file_name": name, "file_name_no_extension": os.path.splitext(name)[0],
               "_abs": os.path.join(base, rel), "_link": link, "_container": container,
               "_is_file_share": bool(az.get("is_file_share", False))}


def _file_factory(config: dict, task: dict, spec: dict) -> Iterator[dict]:
    key = spec["task_filepath"]
    for kw in enumerate_storage_files(config, spec):
        t = _apply(task, (), {k: v for k, v in kw.items() if not k.startswith("_")})
        dest = kw[key]
        # each generated task gets "its" file staged into the working directory
        rf = list(t.get("resource_files") or [])
        rf.append({"file_path": dest, "blob_source": kw["url"]})
        t["resource_files"] = rf
        yield t


def _custom_factory(task: dict, spec: dict) -> Iterator[dict]:
    mod = importlib.import_module(spec["module"], package=spec.get("package"))
    gen = getattr(mod, "generate", None)
    if gen is None:
        raise ValueError(f"custom task factory module '{spec['module']}' has no generate()")
    for args in gen(*(spec.get("input_args") or []), **(spec.get("input_kwargs") or {})):
        if isinstance(args, dict):
            yield _apply(task, (), args)
        else:
            if isinstance(args, (str, bytes)) or not hasattr(args, "__iter__"):
                raise ValueError("custom generate() must yield an iterable (e.g. a tuple) per task")
            yield _apply(task, tuple(args))


def generate_tasks(config: dict, task: dict) -> Iterator[dict]:
    """Expand ``task['task_factory']``; a task without one is yielded unchanged."""
    tf = task.get("task_factory")
    if not tf:
        yield copy.deepcopy(task)
        return
    if tf.get("parametric_sweep") is not None:
        yield from _parametric(task, tf["parametric_sweep"])
    elif tf.get("random") is not None:
        yield from _random_factory(task, tf["random"])
    elif tf.get("repeat") is not None:
        for _ in range(int(tf["repeat"])):
            yield _tmpl(task)
    elif tf.get("file") is not None:
        yield from _file_factory(config, task, tf["file"])
    elif tf.get("custom") is not None:
        yield from _custom_factory(task, tf["custom"])
    else:
        raise ValueError("task_factory needs one of parametric_sweep, random, repeat, file, custom")
