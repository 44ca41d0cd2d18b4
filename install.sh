#!/usr/bin/env bash
# Offline installer: builds the native sm_100a runtime in-tree and puts `shipyard` on PATH via a venv-free symlink.
# (reference: /root/reference/install.sh creates a virtualenv and pip-installs Azure SDKs — nothing to fetch here)
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PY="${PYTHON:-python3}"
"$PY" - <<'PYEOF'
import sys
missing = [m for m in ("click", "yaml", "torch") if __import__("importlib").util.find_spec(m) is None]
if missing:
    sys.exit("missing python modules: " + ", ".join(missing))
PYEOF
command -v nvcc >/dev/null || { echo "nvcc not found (need CUDA >= 12.8 for sm_100a)"; exit 1; }
"$PY" "$here/native/build.py"
bindir="${1:-$HOME/.local/bin}"
mkdir -p "$bindir"
ln -sf "$here/shipyard" "$bindir/shipyard"
echo "installed: $bindir/shipyard  (native libs in $here/batch_shipyard_b200/_native)"
