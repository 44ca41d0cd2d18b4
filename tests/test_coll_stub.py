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


def test_tuning_knobs_roundtrip_and_clamp():
    """Every algorithm-selection knob documented in docs/native-libraries.md is settable at run time, readable back, and clamped to what
    the heap layout can hold (SY_LL_MAX_PAYLOAD 16 KB, SY_LM_MAX_PAYLOAD 256 KB, SY_OS_SLOT 1 MB)."""
    import uuid
    from batch_shipyard_b200.ops.coll import Communicator, CollError
    comm = Communicator(0, 1, "tune" + uuid.uuid4().hex[:8], None, heap_bytes=64 << 20)
    try:
        defaults = {"ll_max_bytes": 16384, "lm_max_bytes": 256 << 10, "oneshot_max_bytes": 256 << 10, "mailbox_max_bytes": 1 << 20,
                    "ag_p2p_min_bytes": 16 << 20, "bcast_sag_min_bytes": 8 << 20, "nvls_min_world": 4, "max_blocks": 128}
        for k, v in defaults.items():
            assert comm.get_tuning(k) == v, (k, comm.get_tuning(k))
        comm.set_tuning(lm_max_bytes=0, ag_p2p_min_bytes=1 << 30, ll_max_bytes=4096)
        assert (comm.get_tuning("lm_max_bytes"), comm.get_tuning("ag_p2p_min_bytes"), comm.get_tuning("ll_max_bytes")) == (0, 1 << 30, 4096)
        comm.set_tuning(lm_max_bytes=1 << 30, ll_max_bytes=1 << 30, mailbox_max_bytes=1 << 30, max_blocks=100000)
        assert comm.get_tuning("lm_max_bytes") == 256 << 10 and comm.get_tuning("ll_max_bytes") == 16384
        assert comm.get_tuning("mailbox_max_bytes") == 1 << 20 and comm.get_tuning("max_blocks") == 256
        with pytest.raises(CollError):
            comm.set_tuning(no_such_knob=1)
    finally:
        comm.close()
