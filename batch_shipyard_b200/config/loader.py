"""Config discovery, loading, validation and merge.

Behaviour parity with the reference CLI context
(/root/reference/shipyard.py:389-575): each config kind may be given by an
explicit path (flag or ``SHIPYARD_<KIND>_CONF`` env var) or found in
``--configdir`` (env ``SHIPYARD_CONFIGDIR``, default cwd) as
``<name>.yaml|.yml|.json``; every file is validated against its schema and
then deep-merged (mappings recurse, lists replace) into ONE dict in the fixed
order credentials -> config -> pool -> jobs -> fs -> monitor -> federation ->
slurm.  Hidden keys ``_verbose/_auto_confirm/_raw`` carry CLI switches.
"""
from __future__ import annotations

import json
import os
from typing import Optional

import yaml

from ..utils.util import merge_dict
from .schema import ConfigType, ValidationError, validate

_KIND_FILES = {
    ConfigType.Credentials: "credentials",
    ConfigType.Global: "config",
    ConfigType.Pool: "pool",
    ConfigType.Jobs: "jobs",
    ConfigType.RemoteFS: "fs",
    ConfigType.Monitor: "monitor",
    ConfigType.Federation: "federation",
    ConfigType.Slurm: "slurm",
}
_ENV = {
    ConfigType.Credentials: "SHIPYARD_CREDENTIALS_CONF",
    ConfigType.Global: "SHIPYARD_CONFIG_CONF",
    ConfigType.Pool: "SHIPYARD_POOL_CONF",
    ConfigType.Jobs: "SHIPYARD_JOBS_CONF",
    ConfigType.RemoteFS: "SHIPYARD_FS_CONF",
    ConfigType.Monitor: "SHIPYARD_MONITOR_CONF",
    ConfigType.Federation: "SHIPYARD_FEDERATION_CONF",
    ConfigType.Slurm: "SHIPYARD_SLURM_CONF",
}
MERGE_ORDER = [ConfigType.Credentials, ConfigType.Global, ConfigType.Pool, ConfigType.Jobs, ConfigType.RemoteFS,
               ConfigType.Monitor, ConfigType.Federation, ConfigType.Slurm]


class ConfigError(Exception):
    pass


def load_file(path: str) -> dict:
    """Parse YAML (or JSON: a YAML subset, but .json gets the strict parser for clear errors)."""
    with open(path, "r") as f:
        text = f.read()
    try:
        data = json.loads(text) if path.endswith(".json") else yaml.safe_load(text)
    except Exception as e:  # noqa: BLE001 - surface parser errors uniformly
        raise ConfigError(f"{path}: cannot parse: {e}") from e
    if data is None:
        return {}
    if not isinstance(data, dict):
        raise ConfigError(f"{path}: top level must be a mapping")
    return data


def form_conf_path(explicit: Optional[str], configdir: Optional[str], kind: ConfigType) -> Optional[str]:
    """explicit flag > env var > configdir/<name>.{yaml,yml,json} > ./<name>.*"""
    if explicit:
        return explicit
    env = os.environ.get(_ENV[kind])
    if env:
        return env
    base = configdir or os.environ.get("SHIPYARD_CONFIGDIR") or "."
    for ext in (".yaml", ".yml", ".json"):
        p = os.path.join(base, _KIND_FILES[kind] + ext)
        if os.path.isfile(p):
            return p
    return None


def normalize(kind: ConfigType, data):
    """Rewrite spellings that older releases of the reference accepted (and some of its own recipes still use) into the current
    form before validation.  Today: ``auto_scratch: true`` on a job (recipes/OpenFOAM-*) -> ``{setup: dependency}``."""
    if kind is ConfigType.Jobs and isinstance(data, dict):
        jobs = data.get("job_specifications")
        for job in jobs if isinstance(jobs, list) else []:            # anything else is left for the validator to report
            if isinstance(job, dict) and job.get("auto_scratch") is True:
                job["auto_scratch"] = {"setup": "dependency"}
            elif isinstance(job, dict) and job.get("auto_scratch") is False:
                job.pop("auto_scratch")
    return data


def load_configs(paths: Optional[dict] = None, configdir: Optional[str] = None,
                 required: tuple = (), skip: tuple = (), verbose: bool = False, auto_confirm: bool = False,
                 raw: bool = False) -> dict:
    """Discover, validate and merge.  ``paths`` maps ConfigType -> explicit path."""
    paths = paths or {}
    merged: dict = {}
    found = {}
    for kind in MERGE_ORDER:
        if kind in skip:
            continue
        p = form_conf_path(paths.get(kind), configdir, kind)
        if p is None:
            if kind in required:
                raise ConfigError(f"{_KIND_FILES[kind]} config is required but was not found "
                                  f"(flag, ${_ENV[kind]}, or --configdir)")
            continue
        if not os.path.isfile(p):
            raise ConfigError(f"{_KIND_FILES[kind]} config '{p}' does not exist")
        data = normalize(kind, load_file(p))
        try:
            validate(kind, data, source=p)
        except ValidationError as e:
            raise ConfigError("invalid configuration:\n  " + "\n  ".join(f"{p}: {m}" for m in e.errors)) from e
        found[kind] = p
        merged = merge_dict(merged, data)
    merged["_verbose"], merged["_auto_confirm"], merged["_raw"] = verbose, auto_confirm, raw
    merged["_config_files"] = {k.value: v for k, v in found.items()}
    return merged


def dump_config(config: dict) -> str:
    """--show-config output: merged config minus hidden keys, secrets masked."""
    def scrub(x, key=""):
        if isinstance(x, dict):
            return {k: scrub(v, str(k)) for k, v in x.items() if not str(k).startswith("_")}
        if isinstance(x, list):
            return [scrub(v, key) for v in x]
        if isinstance(x, str) and any(s in key.lower() for s in ("password", "key", "passphrase", "secret", "token")) \
                and not key.endswith("_id"):
            return "***"
        return x
    return yaml.safe_dump(scrub(config), default_flow_style=False, sort_keys=True)
