"""Launch N rank processes of a worker script and collect their output."""
import os
import subprocess
import sys
import uuid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_ranks(script: str, world: int, extra=(), gpu: bool = False, timeout: int = 300, env=None):
    session = uuid.uuid4().hex[:12]
    procs = []
    e = dict(os.environ)
    e.update(env or {})
    e.setdefault("SHIPYARD_COLL_TIMEOUT_MS", "15000")
    for r in range(world):
        cmd = [sys.executable, os.path.join(ROOT, "tests", script), "--rank", str(r), "--world", str(world),
               "--session", session, "--device", str(r if gpu else -1)] + list(extra)
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=e, cwd=ROOT))
    outs, ok = [], True
    for r, p in enumerate(procs):
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
            out += "\n<<timeout>>"
        outs.append(out)
        ok = ok and p.returncode == 0 and " OK" in out
    return ok, outs
