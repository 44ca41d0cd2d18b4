"""Strict configuration validation with a compact, purpose-built schema notation.

The reference validates every config file with pykwalify against
``schemas/*.yaml`` (/root/reference/convoy/validator.py:52-125).  pykwalify is
not available here and copying those schema files is off the table, so the
same key/type/enum constraints are re-expressed in our own notation, parsed
from ``schemas/<type>.yaml`` next to this file:

    key!: str                 required scalar          key: int:0..100     ranged int
    key: enum:a|b|c           enumeration              key: int|str        union
    key: [str]                list of str              key: {nested map}   strict mapping
    "*": <spec>               arbitrary keys -> spec   key: any            anything

``null`` is accepted for any optional key (treated as absent).  Unknown keys are
errors — same strictness as the reference.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Optional

import yaml


class ConfigType(Enum):
    Credentials = "credentials"
    Global = "config"
    Pool = "pool"
    Jobs = "jobs"
    RemoteFS = "fs"
    Monitor = "monitor"
    Federation = "federation"
    Slurm = "slurm"


class ValidationError(Exception):
    def __init__(self, errors: list[str], source: str = ""):
        self.errors, self.source = errors, source
        super().__init__((f"{source}: " if source else "") + "; ".join(errors))


@dataclass
class Node:
    kind: str                                   # scalar | enum | seq | map | any | union
    types: tuple = ()                           # scalar python type names
    enum: tuple = ()
    lo: Optional[float] = None
    hi: Optional[float] = None
    item: Optional["Node"] = None               # seq item
    keys: dict = field(default_factory=dict)    # map: name -> (Node, required)
    wildcard: Optional["Node"] = None           # map: "*" spec
    alts: tuple = ()                            # union


_SCALARS = {"str": (str,), "int": (int,), "bool": (bool,), "num": (int, float), "float": (int, float)}


def _parse_scalar(spec: str) -> Node:
    spec = spec.strip()
    if spec == "any":
        return Node("any")
    if spec.startswith("enum:"):
        return Node("enum", enum=tuple(v.strip() for v in spec[5:].split("|")))
    if "|" in spec:
        return Node("union", alts=tuple(_parse_scalar(s) for s in spec.split("|")))
    m = re.match(r"^(int|num|float):(-?[0-9.]*)\.\.(-?[0-9.]*)$", spec)
    if m:
        return Node("scalar", types=_SCALARS[m.group(1)], lo=float(m.group(2)) if m.group(2) else None,
                    hi=float(m.group(3)) if m.group(3) else None)
    if spec in _SCALARS:
        return Node("scalar", types=_SCALARS[spec])
    raise ValueError(f"bad schema type '{spec}'")


def compile_schema(raw: Any) -> Node:
    if isinstance(raw, str):
        return _parse_scalar(raw)
    if isinstance(raw, list):
        if len(raw) != 1:
            raise ValueError("a list spec has exactly one item spec")
        return Node("seq", item=compile_schema(raw[0]))
    if isinstance(raw, dict):
        n = Node("map")
        for k, v in raw.items():
            k = str(k)
            if k == "*":
                n.wildcard = compile_schema(v)
                continue
            req = k.endswith("!")
            n.keys[k[:-1] if req else k] = (compile_schema(v), req)
        return n
    raise ValueError(f"bad schema node {raw!r}")


def _check(node: Node, val: Any, path: str, errs: list[str]) -> None:
    if node.kind == "any" or val is None:
        return
    if node.kind == "scalar":
        ok = isinstance(val, node.types) and not (isinstance(val, bool) and bool not in node.types)
        if not ok and str in node.types and isinstance(val, (int, float)) and not isinstance(val, bool):
            ok = True   # YAML turns unquoted digits into numbers; strings accept them
        if not ok:
            errs.append(f"{path}: expected {'/'.join(t.__name__ for t in node.types)}, got {type(val).__name__}")
            return
        if node.lo is not None and val < node.lo:
            errs.append(f"{path}: {val} < minimum {node.lo:g}")
        if node.hi is not None and val > node.hi:
            errs.append(f"{path}: {val} > maximum {node.hi:g}")
    elif node.kind == "enum":
        if str(val) not in node.enum and not (isinstance(val, bool) and str(val).lower() in node.enum):
            errs.append(f"{path}: '{val}' not one of {list(node.enum)}")
    elif node.kind == "union":
        for alt in node.alts:
            sub: list[str] = []
            _check(alt, val, path, sub)
            if not sub:
                return
        errs.append(f"{path}: value {val!r} matches none of the allowed types")
    elif node.kind == "seq":
        if not isinstance(val, list):
            errs.append(f"{path}: expected a list, got {type(val).__name__}")
            return
        for i, v in enumerate(val):
            if v is None and node.item.kind != "any":
                errs.append(f"{path}[{i}]: empty list item")          # `- ` with nothing after it: never meaningful, crashes consumers
                continue
            _check(node.item, v, f"{path}[{i}]", errs)
    elif node.kind == "map":
        if not isinstance(val, dict):
            errs.append(f"{path}: expected a mapping, got {type(val).__name__}")
            return
        for k, (sub, req) in node.keys.items():
            if req and (k not in val or val[k] is None):
                errs.append(f"{path}.{k}: required key missing")
        for k, v in val.items():
            ks = str(k)
            if ks in node.keys:
                _check(node.keys[ks][0], v, f"{path}.{ks}", errs)
            elif node.wildcard is not None:
                if v is None and node.wildcard.kind == "map":
                    errs.append(f"{path}.{ks}: empty entry (a user-named entry needs its settings)")
                    continue
                _check(node.wildcard, v, f"{path}.{ks}", errs)
            else:
                errs.append(f"{path}.{ks}: unknown key")


_CACHE: dict[ConfigType, Node] = {}


def schema_path(ct: ConfigType) -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "schemas", ct.value + ".yaml")


def load_schema(ct: ConfigType) -> Node:
    if ct not in _CACHE:
        with open(schema_path(ct)) as f:
            _CACHE[ct] = compile_schema(yaml.safe_load(f))
    return _CACHE[ct]


def validate(ct: ConfigType, data: Any, source: str = "") -> None:
    """Raise ValidationError listing every problem (unknown key, bad type, bad enum, missing key)."""
    errs: list[str] = []
    if data is None:
        data = {}
    _check(load_schema(ct), data, "$", errs)
    if errs:
        raise ValidationError(errs, source)


def validate_config(ct: ConfigType, path: str) -> dict:
    """Load + validate a YAML/JSON config file; returns the parsed mapping."""
    from .loader import load_file
    data = load_file(path)
    validate(ct, data, source=path)
    return data
