"""Executable model of the parity-double-buffered mailbox protocols of native/coll/kernels.cu (k_lm_k, k_ll, k_mailbox_k, k_oneshot).

What the kernels rely on: every rank numbers its mailbox operations (a device-resident sequence), operation k uses slot parity k & 1,
there is NO barrier between operations, and a writer never checks whether the receiver has consumed the previous contents of a slot.
That is safe only if finishing operation k + 1 implies that every peer has finished reading operation k (then nobody can be two
operations ahead of a reader).  "Every rank hears from every rank in every operation" gives exactly that — which is why the broadcast
modes send token lines / flags all-to-all although only the root has payload.

The model runs W ranks as step machines over shared slot arrays under adversarial (random, seeded) interleavings and checks
  * no lost line: a reader waiting for sequence s never finds a NEWER sequence in the slot it polls (its line was overwritten);
  * no deadlock: all ranks finish every program;
  * every rank received exactly the payload the writer sent in that operation.
It also shows that the root-only-flag broadcast (the round-1 k_mailbox_k) violates the first property.
"""
import random

import pytest


class Violation(Exception):
    pass


class Rank:
    """One rank executing a list of operations; each call to step() performs ONE atomic action (one remote line store, or one poll)."""

    def __init__(self, r, world, program, slots, tokens, all_hear_all_broadcast=True):
        self.r, self.W, self.prog = r, world, program
        self.slots, self.tokens = slots, tokens          # slots[rank][parity][writer] = (seq, payload) ; tokens[rank][parity][writer] = seq
        self.tok_bcast = all_hear_all_broadcast
        self.seq = 0
        self.op = 0
        self.todo = []                                   # remaining atomic actions of the current operation
        self.received = []

    def done(self):
        return self.op >= len(self.prog) and not self.todo

    def _plan(self):
        kind, root = self.prog[self.op]
        s = self.seq + 1
        par = s & 1
        acts = []
        writers = list(range(self.W)) if kind != "bcast" else [root]
        if kind == "bcast" and self.tok_bcast == "root":                              # round-1 k_mailbox_k: non-roots flag the root only
            if self.r != root:
                acts += [("tok", root, par, s)]
        elif kind == "bcast" and self.tok_bcast:
            acts += [("tok", p, par, s) for p in range(self.W)]                       # enter: token to every rank
        if self.r in writers:
            acts += [("put", p, par, s, (self.op, self.r)) for p in range(self.W)]   # payload line to every rank (self included)
        random.shuffle(acts)                                                        # stores of one thread block are unordered across peers
        acts += [("get", w, par, s) for w in writers]                                 # then poll my own slots
        if kind == "bcast" and self.tok_bcast == "root":
            if self.r == root:
                acts += [("gettok", w, par, s) for w in range(self.W) if w != root]   # only the root waits for the others
        elif kind == "bcast" and self.tok_bcast:
            acts += [("gettok", w, par, s) for w in range(self.W)]                    # leave only after every token has arrived
        self.todo = acts

    def step(self):
        """Returns True if progress was made."""
        if self.done():
            return False
        if not self.todo:
            self._plan()
        a = self.todo[0]
        if a[0] == "put":
            _, p, par, s, payload = a
            self.slots[p][par][self.r] = (s, payload)
        elif a[0] == "tok":
            _, p, par, s = a
            self.tokens[p][par][self.r] = s
        elif a[0] == "get":
            _, w, par, s = a
            have = self.slots[self.r][par][w]
            if have[0] > s:
                raise Violation(f"rank {self.r} waits for seq {s} from {w} but the slot already holds seq {have[0]} (line overwritten)")
            if have[0] < s:
                return False                                                         # not there yet: spin
            self.received.append((self.op, w, have[1]))
        elif a[0] == "gettok":
            _, w, par, s = a
            have = self.tokens[self.r][par][w]
            if have > s:
                raise Violation(f"rank {self.r} waits for token {s} from {w} but sees {have}")
            if have < s:
                return False
        self.todo.pop(0)
        if not self.todo:
            self.seq += 1
            self.op += 1
        return True


def run(world, program, seed, all_hear_all_broadcast=True, bias=None):
    rnd = random.Random(seed)
    random.seed(seed)
    slots = [[[(0, None) for _ in range(world)] for _ in range(2)] for _ in range(world)]
    tokens = [[[0 for _ in range(world)] for _ in range(2)] for _ in range(world)]
    ranks = [Rank(r, world, program, slots, tokens, all_hear_all_broadcast) for r in range(world)]
    idle = 0
    while not all(k.done() for k in ranks):
        # adversarial scheduler: `bias` makes one rank slow (it is picked rarely), the classic way to expose run-ahead
        weights = [(0.02 if (bias is not None and r == bias) else 1.0) for r in range(world)]
        k = rnd.choices(ranks, weights)[0]
        if k.step():
            idle = 0
        else:
            idle += 1
            if idle > 20000 * world:
                raise Violation("deadlock: no rank can make progress")
    for k in ranks:
        for op, (kind, root) in enumerate(program):
            got = sorted((w, p) for (o, w, p) in k.received if o == op)
            want = sorted((w, (op, w)) for w in (range(world) if kind != "bcast" else [root]))
            assert got == want, (k.r, op, got, want)


PROGRAMS = [
    [("allgather", 0)] * 6,
    [("allgather", 0), ("bcast", 0), ("allgather", 0), ("bcast", 1), ("bcast", 1), ("allgather", 0)],
    [("bcast", 0)] * 5,
    [("bcast", 2), ("allgather", 0), ("allgather", 0), ("bcast", 0), ("allgather", 0), ("bcast", 1), ("allgather", 0)],
]


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("pi", range(len(PROGRAMS)))
def test_all_hear_all_operations_never_lose_a_line(world, pi):
    prog = [(k, r % world) for k, r in PROGRAMS[pi]]
    for seed in range(40):
        run(world, prog, seed)
        run(world, prog, 1000 + seed, bias=seed % world)          # one rank almost starved


def test_root_only_flag_broadcast_can_overwrite_an_unread_line():
    """The protocol k_mailbox_k used in round 1 (in a broadcast the non-roots flag the ROOT and only the root waits for them): with 3 ranks
    a non-root that finishes a broadcast early enters the next all-gather and overwrites the slot a starved non-root has not read yet.
    Having no tokens at all is of course just as unsafe."""
    prog = [("allgather", 0), ("bcast", 0), ("allgather", 0)]
    for mode in ("root", False):
        found = 0
        for seed in range(300):
            try:
                run(3, prog, seed, all_hear_all_broadcast=mode, bias=2)
            except Violation:
                found += 1
        assert found > 0, f"expected the run-ahead window to show up under a starved reader (mode {mode})"
    for seed in range(300):                                         # and the token version closes it under the same schedules
        run(3, prog, seed, all_hear_all_broadcast=True, bias=2)
