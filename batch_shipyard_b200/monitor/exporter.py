"""Prometheus exporter + target discovery for the box.

Replaces node_exporter/cAdvisor on every VM plus the heimdall file-SD daemon
(/root/reference/heimdall/heimdall.py:292-427: poll the monitoring table, list nodes per
pool, write ``file_sd`` JSON only when its sha256 changes).  One exporter serves, in
Prometheus text format: per-GPU utilisation / memory / power / clocks (NVML through
``nvidia-smi`` when present), pool node-state histograms, task counts per job, staged-bytes
counters and timing events — everything `pool stats` / `jobs stats` shows.
"""
from __future__ import annotations

import hashlib
import http.server
import json
import os
import shutil
import subprocess
import threading
import time
from typing import Optional

from ..backend.local import LocalBackend

_GPU_Q = "index,utilization.gpu,memory.used,memory.total,power.draw,clocks.sm,temperature.gpu"


def gpu_metrics() -> list[dict]:
    if not shutil.which("nvidia-smi"):
        return []
    try:
        out = subprocess.run(["nvidia-smi", f"--query-gpu={_GPU_Q}", "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True, timeout=10).stdout
    except Exception:  # noqa: BLE001
        return []
    rows = []
    for line in out.strip().splitlines():
        c = [x.strip() for x in line.split(",")]
        try:
            rows.append({"index": int(c[0]), "util": float(c[1]), "mem_used": float(c[2]) * (1 << 20), "mem_total": float(c[3]) * (1 << 20),
                         "power": float(c[4]), "sm_mhz": float(c[5]), "temp": float(c[6])})
        except (ValueError, IndexError):
            continue
    return rows


def parse_nvlink_counters(text: str) -> list[dict]:
    """`nvidia-smi nvlink -gt d` -> [{gpu, link, tx_bytes, rx_bytes}] (KiB counters; tolerant of layout differences)."""
    import re
    rows: dict = {}
    gpu = None
    for line in text.splitlines():
        mg = re.match(r"\s*GPU\s+(\d+)\s*:", line)
        if mg:
            gpu = int(mg.group(1)); continue
        ml = re.match(r"\s*Link\s+(\d+)\s*:\s*(?:Data\s+)?(Tx|Rx)\s*:\s*([0-9.]+)\s*(KiB|MiB|GiB|B)?", line, re.I)
        if ml and gpu is not None:
            mult = {"b": 1, "kib": 1 << 10, "mib": 1 << 20, "gib": 1 << 30}[(ml.group(4) or "KiB").lower()]
            r = rows.setdefault((gpu, int(ml.group(1))), {"gpu": gpu, "link": int(ml.group(1)), "tx_bytes": 0.0, "rx_bytes": 0.0})
            r["tx_bytes" if ml.group(2).lower() == "tx" else "rx_bytes"] = float(ml.group(3)) * mult
    return [rows[k] for k in sorted(rows)]


def nvlink_metrics() -> list[dict]:
    if not shutil.which("nvidia-smi"):
        return []
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10).stdout
    except Exception:  # noqa: BLE001
        return []
    return parse_nvlink_counters(out)


def render_metrics(b: LocalBackend, pools: Optional[list] = None) -> str:
    L = []

    def m(name, help_, typ, samples):
        L.append(f"# HELP {name} {help_}")
        L.append(f"# TYPE {name} {typ}")
        for labels, v in samples:
            ls = ",".join(f'{k}="{val}"' for k, val in labels.items())
            L.append(f"{name}{{{ls}}} {v}" if ls else f"{name} {v}")

    g = gpu_metrics()
    m("shipyard_gpu_utilization_percent", "GPU utilisation", "gauge", [({"gpu": r["index"]}, r["util"]) for r in g])
    m("shipyard_gpu_memory_used_bytes", "GPU memory in use", "gauge", [({"gpu": r["index"]}, r["mem_used"]) for r in g])
    m("shipyard_gpu_power_watts", "GPU power draw", "gauge", [({"gpu": r["index"]}, r["power"]) for r in g])
    m("shipyard_gpu_sm_clock_mhz", "GPU SM clock", "gauge", [({"gpu": r["index"]}, r["sm_mhz"]) for r in g])
    nv = nvlink_metrics()
    m("shipyard_nvlink_tx_bytes_total", "NVLink bytes transmitted per GPU link", "counter", [({"gpu": r["gpu"], "link": r["link"]}, int(r["tx_bytes"])) for r in nv])
    m("shipyard_nvlink_rx_bytes_total", "NVLink bytes received per GPU link", "counter", [({"gpu": r["gpu"], "link": r["link"]}, int(r["rx_bytes"])) for r in nv])
    node_s, task_s, slot_s = [], [], []
    for p in b.list_pools():
        if pools and p["id"] not in pools:
            continue
        hist: dict = {}
        for n in b.list_nodes(p["id"]):
            hist[n["state"]] = hist.get(n["state"], 0) + 1
        node_s += [({"pool": p["id"], "state": s}, c) for s, c in sorted(hist.items())]
        st = b.pool_stats(p["id"])
        slot_s.append(({"pool": p["id"]}, st["slot_utilization_pct"]))
        for j in b.list_jobs(p["id"]):
            for s, c in b.count_tasks(j["id"]).items():
                task_s.append(({"pool": p["id"], "job": j["id"], "state": s}, c))
    m("shipyard_pool_nodes", "Nodes per pool and state", "gauge", node_s)
    m("shipyard_pool_slot_utilization_percent", "Task slot utilisation", "gauge", slot_s)
    m("shipyard_job_tasks", "Tasks per job and state", "gauge", task_s)
    ev: dict = {}
    for e in b.store.events():
        k = (e["pool"] or "", f"{e['source']}:{e['event']}")
        ev[k] = ev.get(k, 0) + 1
    m("shipyard_timing_events_total", "Timing events recorded", "counter", [({"pool": k[0], "event": k[1]}, c) for k, c in sorted(ev.items())])
    # per-collective latency histograms written by traced communicators (SHIPYARD_TRACE, ops/coll.py)
    agg: dict = {}
    mdir = os.path.join(b.root, "metrics")
    if os.path.isdir(mdir):
        for fn in sorted(os.listdir(mdir)):
            if not (fn.startswith("coll-") and fn.endswith(".json")):
                continue
            try:
                with open(os.path.join(mdir, fn)) as f:
                    doc = json.load(f)
            except (OSError, ValueError):
                continue
            for op, h in (doc.get("ops") or {}).items():
                a = agg.setdefault(op, {"count": 0, "sum_us": 0.0, "buckets": [0] * len(h["buckets"]), "edges": doc.get("buckets_us") or []})
                a["count"] += h["count"]; a["sum_us"] += h["sum_us"]
                a["buckets"] = [x + y for x, y in zip(a["buckets"], h["buckets"])]
    bucket_s, count_s, sum_s = [], [], []
    for op, a in sorted(agg.items()):
        run = 0
        for edge, c in zip(list(a["edges"]) + ["+Inf"], a["buckets"]):
            run += c
            bucket_s.append(({"op": op, "le": edge}, run))
        count_s.append(({"op": op}, a["count"])); sum_s.append(({"op": op}, round(a["sum_us"], 2)))
    m("shipyard_collective_latency_us_bucket", "Device-side latency of collective calls (cumulative histogram)", "histogram", bucket_s)
    m("shipyard_collective_latency_us_count", "Collective calls observed", "counter", count_s)
    m("shipyard_collective_latency_us_sum", "Summed device-side latency in microseconds", "counter", sum_s)
    return "\n".join(L) + "\n"


def write_file_sd(b: LocalBackend, path: str, port: int) -> bool:
    """Service discovery file for Prometheus; rewritten only if the content hash changed."""
    targets = []
    for t in b.store.query("monitortarget"):
        targets.append({"targets": [f"127.0.0.1:{port}"], "labels": {"kind": t["_pk"], "id": t["_rk"], "env": "shipyard-b200"}})
    body = json.dumps(targets, indent=1, sort_keys=True)
    new = hashlib.sha256(body.encode()).hexdigest()
    old = None
    if os.path.exists(path):
        with open(path, "rb") as f:
            old = hashlib.sha256(f.read()).hexdigest()
    if new == old:
        return False
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        f.write(body)
    os.replace(tmp, path)
    return True


def serve(state_dir: str, port: int = 9100, polling_interval: float = 15.0, once: bool = False) -> None:
    b = LocalBackend(state_dir=state_dir)
    sd_path = os.path.join(state_dir, "monitor", "file_sd.json")
    os.makedirs(os.path.dirname(sd_path), exist_ok=True)

    class H(http.server.BaseHTTPRequestHandler):
        def do_GET(self):  # noqa: N802
            if self.path.startswith("/metrics"):
                pools = [t["_rk"] for t in b.store.query("monitortarget", "pool")] or None
                body = render_metrics(b, pools).encode()
                self.send_response(200); self.send_header("Content-Type", "text/plain; version=0.0.4")
            elif self.path.startswith("/targets"):
                write_file_sd(b, sd_path, port)
                body = open(sd_path, "rb").read()
                self.send_response(200); self.send_header("Content-Type", "application/json")
            else:
                body = b"shipyard exporter: /metrics /targets\n"
                self.send_response(200)
            self.send_header("Content-Length", str(len(body))); self.end_headers(); self.wfile.write(body)

        def log_message(self, *a):
            pass

    def sd_loop():
        while True:
            write_file_sd(b, sd_path, port)
            if once:
                return
            time.sleep(polling_interval)

    threading.Thread(target=sd_loop, daemon=True).start()
    srv = http.server.ThreadingHTTPServer(("127.0.0.1", port), H)
    srv.serve_forever()


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--state-dir", required=True); ap.add_argument("--port", type=int, default=9100)
    ap.add_argument("--polling-interval", type=float, default=15.0)
    a = ap.parse_args(argv)
    serve(a.state_dir, a.port, a.polling_interval)
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main())
