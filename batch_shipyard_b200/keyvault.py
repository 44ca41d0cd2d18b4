"""Local secret store with the key-vault verbs and ``*_keyvault_secret_id`` substitution.

Parity with /root/reference/convoy/keyvault.py:52-268 (store/fetch credentials as a secret,
list/delete, resolve ``<key>_keyvault_secret_id`` references inside the config).  Secrets are
files (mode 0600) under ``<state>/keyvault``; a secret id is ``https://local.vault/secrets/<name>``
or just ``<name>``.
"""
from __future__ import annotations

import base64
import json
import os
import zlib
from typing import Any

import yaml


def _dir(state_dir: str) -> str:
    d = os.path.join(state_dir, "keyvault")
    os.makedirs(d, exist_ok=True)
    os.chmod(d, 0o700)
    return d


def _name(secret_id: str) -> str:
    return secret_id.rstrip("/").split("/secrets/")[-1].split("/")[0]


def store_secret(state_dir: str, name: str, value: str) -> dict:
    p = os.path.join(_dir(state_dir), name)
    with open(os.open(p, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600), "w") as f:
        f.write(value)
    return {"id": f"https://local.vault/secrets/{name}", "name": name}


def store_credentials(state_dir: str, name: str, creds: dict) -> dict:
    blob = base64.b64encode(zlib.compress(yaml.safe_dump(creds).encode())).decode()
    return store_secret(state_dir, name, blob)


def get_secret(state_dir: str, secret_id: str) -> str:
    p = os.path.join(_dir(state_dir), _name(secret_id))
    if not os.path.exists(p):
        raise KeyError(f"secret {secret_id} not found in the local key vault")
    with open(p) as f:
        return f.read()


def fetch_credentials(state_dir: str, secret_id: str) -> dict:
    return yaml.safe_load(zlib.decompress(base64.b64decode(get_secret(state_dir, secret_id))).decode())


def delete_secret(state_dir: str, name: str) -> dict:
    p = os.path.join(_dir(state_dir), _name(name))
    ok = os.path.exists(p)
    if ok:
        os.remove(p)
    return {"deleted": ok, "name": _name(name)}


def list_secrets(state_dir: str) -> list:
    return [{"id": f"https://local.vault/secrets/{n}", "name": n} for n in sorted(os.listdir(_dir(state_dir)))]


def resolve_secret_ids(state_dir: str, conf: Any) -> Any:
    """Replace ``X_keyvault_secret_id: id`` by ``X: <secret>`` recursively (env-var maps are merged)."""
    if isinstance(conf, list):
        return [resolve_secret_ids(state_dir, c) for c in conf]
    if not isinstance(conf, dict):
        return conf
    out = {}
    for k, v in conf.items():
        if isinstance(k, str) and k.endswith("_keyvault_secret_id") and isinstance(v, str) and v:
            base = k[: -len("_keyvault_secret_id")]
            try:
                secret = get_secret(state_dir, v)
            except KeyError:
                out[k] = v
                continue
            if base == "environment_variables":
                try:
                    extra = json.loads(secret)
                except ValueError:
                    extra = yaml.safe_load(secret)
                merged = dict(extra or {})
                merged.update(out.get(base) or conf.get(base) or {})
                out[base] = merged
            elif out.get(base) is None and conf.get(base) is None:
                out[base] = secret
        else:
            val = resolve_secret_ids(state_dir, v)
            if not (k in out and out[k] is not None and val is None):
                out[k] = val if not (isinstance(out.get(k), dict) and isinstance(val, dict)) else dict(val, **out[k])
    return out
