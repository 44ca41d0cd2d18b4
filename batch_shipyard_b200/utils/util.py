"""General helpers: logging, dict merge, shell wrapping, timedeltas, hashing, subprocess, CIDR.

Capability parity with /root/reference/convoy/util.py (logger setup :86-112,
merge_dict :203, confirm_action :181, wrap_commands_in_shell :347-393, timedelta
parsing :419-458, hashing :461-507, subprocess helpers :519-657, CIDR math :659,
singularity image-name mangling :292-345) — re-implemented, not copied.
"""
from __future__ import annotations

import base64
import copy
import datetime
import hashlib
import ipaddress
import logging
import os
import re
import shlex
import subprocess
import sys
import time
from typing import Any, Iterable, Optional, Sequence

_LOGGERS: dict[str, logging.Logger] = {}


def setup_logger(name: str = "shipyard", verbose: bool = False, stream=None, logfile: Optional[str] = None) -> logging.Logger:
    """One handler per logger (stderr, or ``logfile`` when given); verbose adds origin (module:func:line)."""
    lg = logging.getLogger(name)
    lg.setLevel(logging.DEBUG if verbose else logging.INFO)
    fmt = "%(asctime)s %(levelname)s %(name)s:%(funcName)s:%(lineno)d - %(message)s" if verbose \
        else "%(asctime)s %(levelname)s - %(message)s"
    for h in list(lg.handlers):
        lg.removeHandler(h)
    h = logging.FileHandler(logfile, encoding="utf-8") if logfile else logging.StreamHandler(stream or sys.stderr)
    h.setFormatter(logging.Formatter(fmt))
    lg.addHandler(h)
    lg.propagate = False
    _LOGGERS[name] = lg
    return lg


def get_logger(name: str = "shipyard") -> logging.Logger:
    return _LOGGERS.get(name) or setup_logger(name)


def is_none_or_empty(x: Any) -> bool:
    return x is None or (hasattr(x, "__len__") and len(x) == 0)


def is_not_empty(x: Any) -> bool:
    return not is_none_or_empty(x)


def merge_dict(base: dict, over: dict) -> dict:
    """Deep merge: mappings recurse, everything else (lists included) is replaced by `over`."""
    if not isinstance(base, dict) or not isinstance(over, dict):
        raise ValueError("merge_dict needs two mappings")
    out = copy.deepcopy(base)
    for k, v in over.items():
        if k in out and isinstance(out[k], dict) and isinstance(v, dict):
            out[k] = merge_dict(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def confirm_action(config: dict, msg: Optional[str] = None, allow_auto: bool = True) -> bool:
    """Ask for y/n unless `-y` (config['_auto_confirm'])."""
    if allow_auto and config.get("_auto_confirm", False):
        return True
    prompt = f"Confirm {msg} [y/n]: " if msg else "Confirm [y/n]: "
    while True:
        try:
            a = input(prompt).strip().lower()
        except EOFError:
            return False
        if a in ("y", "yes"):
            return True
        if a in ("n", "no"):
            return False


def wrap_commands_in_shell(commands: Sequence[str], windows: bool = False, wait: bool = True) -> str:
    """Join commands into one fail-fast shell invocation."""
    if windows:
        return "cmd.exe /c \"{}\"".format(" && ".join(commands))
    body = "; ".join(commands)
    if wait:
        body += "; wait"
    return "/bin/bash -c " + shlex.quote("set -e; set -o pipefail; " + body)


def wrap_local_commands_in_shell(commands: Sequence[str]) -> str:
    return wrap_commands_in_shell(commands, wait=False)


_TD_RE = re.compile(r"^(?:(?P<d>\d+)\.)?(?P<h>\d{1,2}):(?P<m>\d{1,2}):(?P<s>\d{1,2})(?:\.(?P<f>\d+))?$")


def convert_string_to_timedelta(value: Optional[str]) -> Optional[datetime.timedelta]:
    """Parse ``[d.]HH:MM:SS[.ffffff]`` (the config surface's duration syntax)."""
    if value is None:
        return None
    if isinstance(value, (int, float)):
        return datetime.timedelta(seconds=value)
    m = _TD_RE.match(str(value).strip())
    if not m:
        raise ValueError(f"'{value}' is not a [d.]HH:MM:SS duration")
    frac = m.group("f")
    return datetime.timedelta(days=int(m.group("d") or 0), hours=int(m.group("h")), minutes=int(m.group("m")),
                              seconds=int(m.group("s")), microseconds=int((frac or "0").ljust(6, "0")[:6]))


def timedelta_to_string(td: datetime.timedelta) -> str:
    total = int(td.total_seconds())
    d, rem = divmod(total, 86400)
    h, rem = divmod(rem, 3600)
    m, s = divmod(rem, 60)
    return (f"{d}." if d else "") + f"{h:02d}:{m:02d}:{s:02d}"


def datetime_utcnow(as_string: bool = False):
    now = datetime.datetime.now(datetime.timezone.utc)
    return now.strftime("%Y%m%dT%H%M%SZ") if as_string else now


def compute_sha256_for_file(path: str, as_base64: bool = False, blocksize: int = 1 << 20) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(blocksize), b""):
            h.update(chunk)
    return base64.b64encode(h.digest()).decode() if as_base64 else h.hexdigest()


def compute_md5_for_file(path: str, as_base64: bool = True, blocksize: int = 1 << 20) -> str:
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(blocksize), b""):
            h.update(chunk)
    return base64.b64encode(h.digest()).decode() if as_base64 else h.hexdigest()


def hash_string(s: str) -> str:
    return hashlib.sha1(s.encode("utf8")).hexdigest()


def hash_federation_id(fid: str) -> str:
    return hashlib.sha1(fid.lower().encode("utf8")).hexdigest()


def base64_encode_string(s) -> str:
    if isinstance(s, str):
        s = s.encode("utf8")
    return base64.b64encode(s).decode("ascii")


def base64_decode_string(s: str) -> str:
    return base64.b64decode(s).decode("utf8")


def subprocess_with_output(cmd, shell: bool = False, cwd: Optional[str] = None, env: Optional[dict] = None,
                           suppress_output: bool = False) -> int:
    if suppress_output:
        return subprocess.call(cmd, shell=shell, cwd=cwd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return subprocess.call(cmd, shell=shell, cwd=cwd, env=env)


def subprocess_capture(cmd, shell: bool = False, cwd: Optional[str] = None, env: Optional[dict] = None,
                       timeout: Optional[float] = None) -> tuple[int, str, str]:
    p = subprocess.run(cmd, shell=shell, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=timeout)
    return p.returncode, p.stdout, p.stderr


def subprocess_nowait(cmd, shell: bool = False, cwd: Optional[str] = None, env: Optional[dict] = None,
                      pipe_stdout: bool = False) -> subprocess.Popen:
    return subprocess.Popen(cmd, shell=shell, cwd=cwd, env=env, stdout=subprocess.PIPE if pipe_stdout else None)


def subprocess_wait_all(procs: Iterable[subprocess.Popen], poll: float = 0.05) -> list[int]:
    procs = list(procs)
    rcs: list[Optional[int]] = [None] * len(procs)
    while any(rc is None for rc in rcs):
        for i, p in enumerate(procs):
            if rcs[i] is None:
                rcs[i] = p.poll()
        if any(rc is None for rc in rcs):
            time.sleep(poll)
    return [int(rc) for rc in rcs]


def subprocess_wait_any(procs: Sequence[subprocess.Popen], poll: float = 0.05) -> tuple[int, int]:
    while True:
        for i, p in enumerate(procs):
            rc = p.poll()
            if rc is not None:
                return i, rc
        time.sleep(poll)


def ip_from_address_prefix(cidr: str, start_offset: int = 4, max_hosts: Optional[int] = None) -> list[str]:
    """Usable addresses of a CIDR block (first `start_offset` are reserved, as on a cloud vnet)."""
    net = ipaddress.ip_network(cidr, strict=False)
    hosts = [str(h) for i, h in enumerate(net.hosts()) if i + 1 >= start_offset]
    return hosts[:max_hosts] if max_hosts else hosts


def explode_arm_subnet_id(arm_id: str) -> tuple[str, str, str, str]:
    """/subscriptions/S/resourceGroups/RG/providers/Microsoft.Network/virtualNetworks/V/subnets/N -> (S, RG, V, N)."""
    t = arm_id.strip("/").split("/")
    if len(t) < 10 or t[0].lower() != "subscriptions":
        raise ValueError(f"'{arm_id}' is not a subnet resource id")
    return t[1], t[3], t[7], t[9]


_SING_PREFIXES = ("shub://", "docker://", "library://", "oras://", "http://", "https://")


def singularity_image_name_on_disk(name: str) -> str:
    """File name a singularity pull of `name` leaves in the image cache."""
    docker = name.startswith("docker://")
    for pfx in _SING_PREFIXES:
        if name.startswith(pfx):
            name = name[len(pfx):]
            break
    # digests / tags become part of the file name
    name = name.replace("@", "_").replace("/", "-")
    if ":" in name:
        base, tag = name.rsplit(":", 1)
    else:
        base, tag = name, ("latest" if docker else "master")
    return f"{base}_{tag}.sif"


def normalize_docker_image_name_for_job(job_id: str, image: str) -> str:
    """Container name for a multi-instance coordination container: <jobid>-<image> with safe chars."""
    return re.sub(r"[^a-zA-Z0-9_.-]", "-", f"{job_id}-{image}")


def parse_size_to_mb(value) -> Optional[float]:
    """'512m' / '2g' / '1t' / '1024k' / bytes-int -> megabytes."""
    if value is None:
        return None
    if isinstance(value, (int, float)):
        return float(value) / (1 << 20)
    m = re.match(r"^\s*([0-9.]+)\s*([bkmgtBKMGT]?)", str(value))
    if not m:
        raise ValueError(f"cannot parse size '{value}'")
    num = float(m.group(1))
    mult = {"": 1.0 / (1 << 20), "b": 1.0 / (1 << 20), "k": 1.0 / 1024, "m": 1.0, "g": 1024.0, "t": 1024.0 * 1024}
    return num * mult[m.group(2).lower()]


def expand_env(value: str, env: Optional[dict] = None) -> str:
    """Expand $VAR / ${VAR} from `env` (defaults to os.environ); unknown names are left alone."""
    env = os.environ if env is None else env

    def sub(m):
        k = m.group(1) or m.group(2)
        return str(env.get(k, m.group(0)))

    return re.sub(r"\$(?:\{(\w+)\}|(\w+))", sub, value)
