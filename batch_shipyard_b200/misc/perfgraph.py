"""Pool start-up timeline from the perf event log (standalone tool, like the reference's grapher).

The reference records ``(pool, ts, source:event, node, message)`` rows from nodeprep and cascade
(/root/reference/cascade/perf.py:55-82) and a separate script coalesces them into per-node
timelines for gnuplot (/root/reference/cascade/graph.py:169-365).  Here the events live in the state
store (``Store.record_event``); this tool folds start/end pairs into intervals per node and emits
either a text Gantt, JSON, or a gnuplot data+script pair.

    python -m batch_shipyard_b200.misc.perfgraph --pool mypool [--format text|json|gnuplot] [--out prefix]
"""
from __future__ import annotations

import argparse
import json
import os
import re
import sys
from collections import defaultdict
from typing import Optional

from ..state.store import Store

_PAIRS = {"start": "end", "pull-start": "pull-end", "load-start": "load-end", "save-start": "save-end"}


def coalesce(events: list[dict]) -> dict:
    """events -> {node: {"intervals": [(label, t0, t1)], "marks": [(label, t)]}} with times relative to the
    first event of the pool.  ``x-start:<key>`` pairs with ``x-end:<key>`` (key = image / resource name)."""
    if not events:
        return {}
    t_origin = min(e["ts"] for e in events)
    nodes: dict = defaultdict(lambda: {"intervals": [], "marks": []})
    open_: dict = {}
    for e in sorted(events, key=lambda e: e["ts"]):
        node = e["node"] or "-"
        ev, msg = e["event"], (e["message"] or "")
        key = re.split(r"[ ,]", msg)[0] if (msg and ev.startswith(("pull", "load", "save"))) else ""
        label = f'{e["source"]}:{ev}'
        t = e["ts"] - t_origin
        if ev in _PAIRS:
            closer = "gr-done" if (e["source"], ev) == ("cascade", "start") else _PAIRS[ev]
            open_[(node, e["source"], closer, key)] = (f'{e["source"]}:{ev.rsplit("-", 1)[0] if "-" in ev else "run"}' + (f"[{key}]" if key else ""), t)
        elif (node, e["source"], ev, key) in open_:
            name, t0 = open_.pop((node, e["source"], ev, key))
            nodes[node]["intervals"].append((name, t0, t))
        else:
            nodes[node]["marks"].append((label + (f"[{key}]" if key else ""), t))
    for (node, _src, _ev, _key), (name, t0) in open_.items():       # never closed: show as open-ended mark
        nodes[node]["marks"].append((name + ":unfinished", t0))
    return dict(nodes)


def render_text(tl: dict, width: int = 72) -> str:
    if not tl:
        return "(no events)"
    t_max = max([t1 for n in tl.values() for _, _, t1 in n["intervals"]] + [t for n in tl.values() for _, t in n["marks"]] + [1e-9])
    out = []
    for node in sorted(tl):
        out.append(f"{node}")
        for name, t0, t1 in sorted(tl[node]["intervals"], key=lambda x: x[1]):
            a, b = int(t0 / t_max * width), max(int(t1 / t_max * width), int(t0 / t_max * width) + 1)
            out.append(f"  {name[:34]:<34} |{' ' * a}{'#' * (b - a)}{' ' * (width - b)}| {t0:8.3f}s -> {t1:8.3f}s ({t1 - t0:.3f}s)")
        for name, t in sorted(tl[node]["marks"], key=lambda x: x[1]):
            a = min(int(t / t_max * width), width - 1)
            out.append(f"  {name[:34]:<34} |{' ' * a}^{' ' * (width - a - 1)}| {t:8.3f}s")
    return "\n".join(out)


def render_gnuplot(tl: dict, prefix: str) -> tuple[str, str]:
    dat, gp = prefix + ".dat", prefix + ".gp"
    with open(dat, "w") as f:
        for i, node in enumerate(sorted(tl)):
            for name, t0, t1 in tl[node]["intervals"]:
                f.write(f'{i} {t0:.6f} {t1:.6f} "{name}" "{node}"\n')
    with open(gp, "w") as f:
        f.write("set terminal pngcairo size 1400,600\nset output '%s.png'\nset xlabel 'seconds since pool start'\n" % os.path.basename(prefix))
        f.write("set ylabel 'node'\nset yrange [-1:%d]\nset style fill solid 0.6\n" % len(tl))
        f.write("plot '%s' using 2:1:2:3:($1-0.35):($1+0.35) with boxxyerrorbars notitle\n" % os.path.basename(dat))
    return dat, gp


def main(argv: Optional[list] = None) -> int:
    ap = argparse.ArgumentParser(prog="perfgraph")
    ap.add_argument("--pool", default=None)
    ap.add_argument("--state-dir", default=os.environ.get("SHIPYARD_STATE_DIR"))
    ap.add_argument("--format", choices=["text", "json", "gnuplot"], default="text")
    ap.add_argument("--out", default="pool_timeline")
    a = ap.parse_args(argv)
    store = Store(a.state_dir) if a.state_dir else Store()
    tl = coalesce(store.events(a.pool))
    if a.format == "json":
        print(json.dumps(tl, indent=1))
    elif a.format == "gnuplot":
        print("wrote %s %s" % render_gnuplot(tl, a.out))
    else:
        print(render_text(tl))
    return 0


if __name__ == "__main__":
    sys.exit(main())
