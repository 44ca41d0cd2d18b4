"""`shipyard monitor *`: lifecycle of the local monitoring service + monitored-resource registry.

Reference: the client provisions a Prometheus/Grafana VM (/root/reference/convoy/monitor.py:
126-711) and registers pools / storage clusters in a monitoring table
(/root/reference/convoy/storage.py:491-643) that heimdall turns into scrape targets.
Here ``create/start`` launch the exporter process on the box, the registry lives in the
state store, and a ready-to-use ``prometheus.yml`` + Grafana dashboard are written next to it.
"""
from __future__ import annotations

import os
import signal
import subprocess
import sys
import time

from ..config import settings as S


def _dir(b) -> str:
    d = os.path.join(b.root, "monitor")
    os.makedirs(d, exist_ok=True)
    return d


def _alive(pid) -> bool:
    try:
        os.kill(int(pid), 0)
        return True
    except (OSError, TypeError, ValueError):
        return False


def start(b, config: dict) -> dict:
    st = b.store.try_get("service", "monitor", "") or {}
    if _alive(st.get("pid")):
        return dict(status(b), note="already running")
    ms = S.monitoring_services(config) if config.get("monitoring") else S.MonitoringServices()
    port = int(os.environ.get("SHIPYARD_EXPORTER_PORT", "9100"))
    d = _dir(b)
    from . import stack
    stack.write_stack(d, int(ms.prometheus_scrape_interval), port, int(ms.prometheus_port))     # prometheus.yml, compose, grafana provisioning + dashboard
    log = open(os.path.join(d, "exporter.log"), "ab")
    p = subprocess.Popen([sys.executable, "-m", "batch_shipyard_b200.monitor.exporter", "--state-dir", b.root, "--port", str(port),
                          "--polling-interval", str(ms.resource_polling_interval)], stdout=log, stderr=log, stdin=subprocess.DEVNULL,
                         start_new_session=True, cwd=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    log.close()                               # the exporter holds its own copy of the descriptor
    b.store.insert("service", "monitor", "", {"pid": p.pid, "port": port, "state": "running", "started": time.time(),
                                               "prometheus_port": ms.prometheus_port}, replace=True)
    return status(b)


def create(b, config: dict) -> dict:
    return start(b, config)


def stop(b) -> dict:
    st = b.store.try_get("service", "monitor", "")
    if st and _alive(st.get("pid")):
        try:
            os.kill(int(st["pid"]), signal.SIGTERM)
        except OSError:
            pass
    if st:
        b.store.merge("service", "monitor", "", {"state": "suspended", "pid": None})
    return status(b)


def destroy(b) -> dict:
    stop(b)
    b.store.delete("service", "monitor", "")
    b.store.delete("monitortarget")
    return {"destroyed": True}


def status(b) -> dict:
    st = b.store.try_get("service", "monitor", "")
    if not st:
        return {"state": "absent"}
    alive = _alive(st.get("pid"))
    return {"state": "running" if alive else ("suspended" if st.get("state") == "suspended" else "dead"), "pid": st.get("pid"),
            "metrics_url": f"http://127.0.0.1:{st.get('port')}/metrics", "config_dir": _dir(b), "targets": list_targets(b)}


def add_targets(b, pools: list, remote_fs: list) -> dict:
    for p in pools:
        if not b.pool_exists(p):
            raise ValueError(f"pool {p} does not exist")
        b.store.insert("monitortarget", "pool", p, {"added": time.time()}, replace=True)
    for r in remote_fs:
        b.store.insert("monitortarget", "remotefs", r, {"added": time.time()}, replace=True)
    return list_targets(b)


def remove_targets(b, all_: bool, pools: list, remote_fs: list) -> dict:
    if all_:
        b.store.delete("monitortarget")
    for p in pools:
        b.store.delete("monitortarget", "pool", p)
    for r in remote_fs:
        b.store.delete("monitortarget", "remotefs", r)
    return list_targets(b)


def list_targets(b) -> dict:
    return {"pools": [t["_rk"] for t in b.store.query("monitortarget", "pool")],
            "remote_fs": [t["_rk"] for t in b.store.query("monitortarget", "remotefs")]}


def shell(b, command) -> dict:
    if not command:
        return {"login": "local", "cwd": _dir(b)}
    p = subprocess.run(" ".join(command), shell=True, cwd=_dir(b), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return {"exit_code": p.returncode, "output": p.stdout}
