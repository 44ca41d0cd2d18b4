"""GPU topology of the box (node inventory for the local pool).

Wraps ``shipyard-gpuprobe`` (native, CUDA runtime/driver + optional NVML).  On a
CPU-only box the probe reports zero GPUs and pools fall back to virtual CPU slots.
Reference analogue: VM size / RDMA class detection in
/root/reference/scripts/shipyard_nodeprep.sh:380-422 and the GPU/RDMA VM-size
tables in /root/reference/convoy/settings.py:59-144.
"""
from __future__ import annotations

import json
import os
import subprocess
from typing import Optional

_CACHE: Optional[dict] = None


def probe(refresh: bool = False) -> dict:
    global _CACHE
    if _CACHE is not None and not refresh:
        return _CACHE
    from .._build import ensure_built, native_dir
    exe = os.path.join(native_dir(), "shipyard-gpuprobe")
    if not os.path.exists(exe) and not os.environ.get("SHIPYARD_FAKE_GPUS"):
        try:
            ensure_built(["gpuprobe"])                                # built artefacts are not in the history; a checkout probes on first use
        except Exception:  # noqa: BLE001 - no compiler: reported as "probe not built" below
            pass
    info = {"gpus": [], "driver_version": 0, "runtime_version": 0, "error": "probe not built", "nvml": False}
    if os.environ.get("SHIPYARD_FAKE_GPUS"):
        n = int(os.environ["SHIPYARD_FAKE_GPUS"])
        info = {"gpus": [{"index": i, "name": "FAKE B200", "cc": "10.0", "sms": 148, "memory_total": 180 << 30,
                          "memory_free": 180 << 30, "multicast": True, "posix_fd_handles": True,
                          "p2p": [{"peer": j, "access": True, "atomics": True} for j in range(n)]} for i in range(n)],
                "driver_version": 0, "runtime_version": 0, "error": None, "nvml": False, "fake": True}
    elif os.path.exists(exe):
        try:
            out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=60).stdout
            info = json.loads(out)
        except Exception as e:  # noqa: BLE001
            info["error"] = f"probe failed: {e}"
    _CACHE = info
    return info


def probe_gpus(refresh: bool = False) -> list[dict]:
    return list(probe(refresh).get("gpus") or [])


def gpu_count() -> int:
    return len(probe_gpus())


def full_p2p(gpus: Optional[list] = None) -> bool:
    gpus = probe_gpus() if gpus is None else gpus
    return all(all(p.get("access") for p in g.get("p2p", [])) for g in gpus) if gpus else False


def nvls_capable(gpus: Optional[list] = None) -> bool:
    gpus = probe_gpus() if gpus is None else gpus
    return bool(gpus) and all(g.get("multicast") for g in gpus)


def describe() -> dict:
    info = probe()
    gpus = info.get("gpus") or []
    return {"gpu_count": len(gpus), "gpu_name": gpus[0]["name"] if gpus else None, "full_p2p": full_p2p(gpus),
            "nvls_multicast": nvls_capable(gpus), "driver_version": info.get("driver_version"),
            "collective_transport": "nvls" if nvls_capable(gpus) and len(gpus) > 1 else ("p2p" if gpus else "stub"),
            "error": info.get("error")}
