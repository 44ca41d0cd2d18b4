"""Collectives on the host shared-memory stub transport (CPU, world_size 1/2/4)."""
import pytest

from _mp import run_ranks


@pytest.mark.parametrize("world", [1, 2, 4])
def test_stub_collectives(world):
    ok, outs = run_ranks("_coll_worker.py", world, extra=["--quick"], gpu=False, timeout=240)
    assert ok, "\n".join(o[-3000:] for o in outs)
    assert all("transport=stub" in o for o in outs)


def test_flag_protocol_message_passing_litmus():
    """acquire/release litmus on the heap flags — 2000 ping-pong rounds, every payload whole and current."""
    ok, outs = run_ranks("_litmus_worker.py", 2, extra=["--rounds", "2000"], gpu=False, timeout=120)
    assert ok, "\n".join(o[-2000:] for o in outs)
    assert all("message-passing rounds" in o for o in outs)


def test_dropped_flag_ends_in_watchdog_timeout_not_a_hang():
    """SURVEY §5.3 fault injection: rank 0's third signal is dropped; both bounded waits must give up with a timeout error."""
    import time
    t0 = time.time()
    ok, outs = run_ranks("_litmus_worker.py", 2, extra=["--rounds", "10", "--expect-timeout"], gpu=False, timeout=60,
                         env={"SHIPYARD_FAULT_INJECT": "drop_signal:0:3", "SHIPYARD_COLL_TIMEOUT_MS": "1500"})
    assert ok, "\n".join(o[-2000:] for o in outs)
    assert all("timeout as expected" in o for o in outs)
    assert time.time() - t0 < 45
