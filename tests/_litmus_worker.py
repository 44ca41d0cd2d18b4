"""Flag-protocol litmus tests on the symmetric heap (message passing, ping-pong, dropped flag), one rank per process.

Message passing (MP): the producer writes a payload into the peer's heap and then raises the peer's signal with release
semantics; the consumer waits on the signal with acquire semantics and must then see the WHOLE payload of that round —
the ordering every collective kernel of native/coll relies on (payload stores, fence, flag; flag acquire, payload loads).
The payload of round i is the constant i, so a stale or torn read is detectable.  The consumer acknowledges on a second
signal (ping-pong), which bounds run-ahead to one round exactly like the mailbox collectives do.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batch_shipyard_b200.ops import coll  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, required=True)
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--session", required=True)
    ap.add_argument("--device", type=int, default=-1)
    ap.add_argument("--rounds", type=int, default=2000)
    ap.add_argument("--expect-timeout", action="store_true")
    a = ap.parse_args()
    dev_index = a.device if a.device >= 0 else None
    if dev_index is not None:
        torch.cuda.set_device(dev_index)
    comm = coll.Communicator(a.rank, a.world, a.session, dev_index, heap_bytes=64 << 20)
    dev = comm.torch_device
    sync = (lambda: None) if comm.is_stub else (lambda: torch.cuda.synchronize(dev))
    R, W = a.rank, a.world
    assert W == 2
    n = 4096                                            # 16 KB payload: many cache lines, so a torn read would show
    box = comm.alloc(n, torch.int32)                    # my inbox (same offset on the peer)
    box.zero_(); sync()
    comm.barrier(); sync()
    off = comm.heap_offset(box)
    peer = 1 - R
    DATA, ACK = 5, 6
    t0 = time.time()
    try:
        for i in range(1, a.rounds + 1):
            if R == 0:
                payload = torch.full((n,), i, dtype=torch.int32, device=dev)
                comm.put_signal(payload, off, peer, sig=DATA)
                comm.wait_signal(ACK, i)                # consumer has read round i: the inbox may be overwritten
                sync()
            else:
                comm.wait_signal(DATA, i)
                sync()
                got = box.clone()
                if int(got.min()) != i or int(got.max()) != i:
                    raise AssertionError(f"MP litmus violated in round {i}: saw [{int(got.min())}, {int(got.max())}]")
                comm.put_signal(got[:4], off, peer, sig=ACK)    # ack (payload irrelevant: 16 bytes into the producer's inbox)
        comm.check_status()
    except coll.CollError as e:
        if a.expect_timeout and ("timeout" in str(e).lower() or "watchdog" in str(e).lower()):
            print(f"rank {R}: watchdog ended the wait after {time.time() - t0:.2f}s: {e}")
            print(f"rank {R} OK (timeout as expected)")
            return
        raise
    if a.expect_timeout and R == 1:
        # GPU transport: the wait kernel records the timeout in the status word instead of returning an error code
        try:
            comm.check_status()
        except coll.CollError as e:
            print(f"rank {R} OK (timeout as expected): {e}")
            return
        raise AssertionError("the dropped flag went unnoticed")
    print(f"rank {R}: {a.rounds} message-passing rounds, {a.rounds / (time.time() - t0):.0f} rounds/s, transport={'stub' if comm.is_stub else 'gpu'}")
    comm.barrier(); sync()
    print(f"rank {R} OK")


if __name__ == "__main__":
    main()
