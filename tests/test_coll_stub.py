"""Collectives on the host shared-memory stub transport (CPU, world_size 1/2/4)."""
import pytest

from _mp import run_ranks


@pytest.mark.parametrize("world", [1, 2, 4])
def test_stub_collectives(world):
    ok, outs = run_ranks("_coll_worker.py", world, extra=["--quick"], gpu=False, timeout=240)
    assert ok, "\n".join(o[-3000:] for o in outs)
    assert all("transport=stub" in o for o in outs)
