"""Block until the pool's global resources are loaded (job-preparation gate).

Stand-in for /root/reference/scripts/wait_for_images.sh:11-59."""
import argparse
import sys

from ..state.store import Store
from .cascade import wait_for_images


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--state-dir", required=True)
    ap.add_argument("--pool", required=True)
    ap.add_argument("--timeout", type=float, default=600.0)
    a = ap.parse_args(argv)
    ok = wait_for_images(Store(a.state_dir), a.pool, a.timeout)
    if not ok:
        print(f"wait_images: global resources of pool {a.pool} are not all loaded", file=sys.stderr)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
