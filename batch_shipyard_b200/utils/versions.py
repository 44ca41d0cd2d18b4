"""Version metadata checks between the CLI and objects it created earlier.

The reference refuses to add jobs to pools created by an incompatible release
(metadata ``BATCH_SHIPYARD_VERSION`` >= 3.8.0, /root/reference/convoy/batch.py:1337,
5100-5134); same idea, our own version line.
"""
from __future__ import annotations

from .. import __version__

MIN_COMPAT = (0, 1, 0)


def _parse(v: str) -> tuple:
    out = []
    for p in str(v).split(".")[:3]:
        digits = "".join(ch for ch in p if ch.isdigit())
        out.append(int(digits) if digits else 0)
    while len(out) < 3:
        out.append(0)
    return tuple(out)


def check_metadata_compat(metadata: dict, what: str = "pool") -> None:
    v = (metadata or {}).get("BATCH_SHIPYARD_VERSION")
    if v is None:
        raise RuntimeError(f"{what} was not created by shipyard (no version metadata); recreate it")
    if _parse(v) < MIN_COMPAT:
        raise RuntimeError(f"{what} was created by shipyard {v}, older than the minimum compatible "
                           f"{'.'.join(map(str, MIN_COMPAT))} (this is {__version__}); recreate it")
