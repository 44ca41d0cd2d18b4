"""CPU model of the index arithmetic of k_lm_k (native/coll/kernels.cu): the same per-thread line loop — two 4-byte payload words per
16-byte line {w0, flag, w1, flag}, `two` tail handling for odd word counts, writer slots per peer, rank-order sums — executed for every
rank of a world over numpy arrays, and compared with the plain definition of each collective.  The GPU suite runs the real kernel at the
world sizes of the box; this covers world 2..8, ragged sizes and in-place buffers on any machine."""
import numpy as np
import pytest

SY_MAXR = 8


def lm_kernel(world, rank, slots, inp, out, nbytes, mode, dt, scale, seq, nthreads=64):
    """One rank's k_lm_k: `slots[p][parity][writer]` is rank p's line buffer (uint32 [lines, 4]); phases are split (send for all ranks is
    run before any receive by the caller) because the model has no concurrency — the protocol side is tests/test_ll_protocol_model.py."""
    parity, flag = seq & 1, np.uint32(seq)
    words = nbytes // 4
    lines = (words + 1) // 2
    wpr = words
    root = dt if mode == 2 else None

    def send():
        for tid in range(nthreads):
            for ln in range(tid, lines, nthreads):
                two = 2 * ln + 1 < words
                if mode == 2:
                    if rank != root:
                        continue
                    w0, w1 = inp[2 * ln], (inp[2 * ln + 1] if two else np.uint32(0))
                    for j in range(world):
                        p = (rank + j) % world
                        slots[p][parity][root][ln] = (w0, flag, w1, flag)
                    continue
                for j in range(world):
                    p = (rank + j) % world
                    if mode in (1, 3):
                        w0 = inp[p * wpr + 2 * ln]
                        w1 = inp[p * wpr + 2 * ln + 1] if two else np.uint32(0)
                    else:
                        w0 = inp[2 * ln]
                        w1 = inp[2 * ln + 1] if two else np.uint32(0)
                    slots[p][parity][rank][ln] = (w0, flag, w1, flag)

    def recv():
        for tid in range(nthreads):
            for ln in range(tid, lines, nthreads):
                two = 2 * ln + 1 < words
                if mode == 2:
                    v = slots[rank][parity][root][ln]
                    assert v[1] == flag and v[3] == flag
                    out[2 * ln] = v[0]
                    if two:
                        out[2 * ln + 1] = v[2]
                    continue
                v = [slots[rank][parity][r][ln] for r in range(world)]
                assert all(x[1] == flag and x[3] == flag for x in v)
                if mode in (0, 1):
                    for r in range(world):
                        out[r * wpr + 2 * ln] = v[r][0]
                        if two:
                            out[r * wpr + 2 * ln + 1] = v[r][2]
                elif dt == "f32":
                    a0 = np.float32(0); a1 = np.float32(0)
                    for r in range(world):
                        a0 = np.float32(a0 + np.array([v[r][0]], np.uint32).view(np.float32)[0])
                        a1 = np.float32(a1 + np.array([v[r][2]], np.uint32).view(np.float32)[0])
                    out[2 * ln] = np.array([a0 * np.float32(scale)], np.float32).view(np.uint32)[0]
                    if two:
                        out[2 * ln + 1] = np.array([a1 * np.float32(scale)], np.float32).view(np.uint32)[0]
    return send, recv


def run_world(world, nbytes, mode, root=0, inplace=False, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    words = nbytes // 4
    per_in = words * (world if mode in (1, 3) else 1)
    per_out = words * (world if mode in (0, 1) else 1)
    lines = (words + 1) // 2
    slots = [[[np.zeros((lines, 4), np.uint32) for _ in range(SY_MAXR)] for _ in range(2)] for _ in range(world)]
    ins, outs = [], []
    for r in range(world):
        if mode in (3, 4):
            x = rng.integers(-1000, 1000, per_in).astype(np.float32).view(np.uint32)       # exactly summable floats
        else:
            x = rng.integers(0, 2 ** 32, per_in, dtype=np.uint64).astype(np.uint32)
        if inplace and mode == 0:                       # all-gather in place: the input is the rank's own block of the output
            o = np.zeros(per_out, np.uint32)
            o[r * words:(r + 1) * words] = x
            ins.append(o[r * words:(r + 1) * words]); outs.append(o)
        elif inplace and mode in (1, 2, 4):
            ins.append(x); outs.append(x)
        else:
            ins.append(x); outs.append(np.zeros(per_out, np.uint32))
    orig = [np.array(i, copy=True) for i in ins]
    phases = [lm_kernel(world, r, slots, ins[r], outs[r], nbytes, mode, root if mode == 2 else "f32", scale, seq=5) for r in range(world)]
    # a thread reads ALL its input words of a line before writing any output word of that line; model that order globally:
    for send, _ in phases:
        send()
    for _, recv in phases:
        recv()
    return orig, outs


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("nbytes", [4, 8, 12, 28, 1000, 4100])
def test_all_gather_and_all_to_all(world, nbytes):
    words = nbytes // 4
    for inplace in (False, True):
        orig, outs = run_world(world, nbytes, 0, inplace=inplace, seed=world + nbytes)
        want = np.concatenate(orig)
        for r in range(world):
            assert np.array_equal(outs[r], want), ("all_gather", r, inplace)
        orig, outs = run_world(world, nbytes, 1, inplace=inplace, seed=3 * world + nbytes)
        for r in range(world):
            want = np.concatenate([orig[w][r * words:(r + 1) * words] for w in range(world)])
            assert np.array_equal(outs[r], want), ("all_to_all", r, inplace)


@pytest.mark.parametrize("world", [2, 5, 8])
@pytest.mark.parametrize("nbytes", [4, 20, 36, 2052])
def test_broadcast_reduce_scatter_all_reduce(world, nbytes):
    words = nbytes // 4
    for root in (0, world - 1):
        orig, outs = run_world(world, nbytes, 2, root=root, inplace=True, seed=root + nbytes)
        for r in range(world):
            assert np.array_equal(outs[r], orig[root]), ("broadcast", r, root)
    orig, outs = run_world(world, nbytes, 3, seed=7 + nbytes, scale=0.5)
    tot = sum(o.view(np.float32).astype(np.float64) for o in orig) * 0.5
    for r in range(world):
        assert np.array_equal(outs[r].view(np.float32).astype(np.float64), tot[r * words:(r + 1) * words]), ("reduce_scatter", r)
    orig, outs = run_world(world, nbytes, 4, inplace=True, seed=11 + nbytes)
    tot = sum(o.view(np.float32).astype(np.float64) for o in orig)
    for r in range(world):
        assert np.array_equal(outs[r].view(np.float32).astype(np.float64), tot), ("all_reduce", r)
