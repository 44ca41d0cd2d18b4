"""Randomised model check of the mbarrier pipeline protocols of the halo-load kernels (native/gemm/conv_halo.inc, wgrad_halo.inc).

The kernels are warp-specialised: a TMA producer thread, an MMA issuer thread and one or two epilogue warpgroups talk only through
mbarriers (phase parity waits, arrive counts, transaction bytes) and asynchronous completions (TMA `complete_tx`, `tcgen05.commit`).
A protocol mistake — a wrong arrive count, a phase that advances twice before its waiter looks, a slot refilled before its readers
are done — shows up on hardware as a hang (the watchdog trap) or as silent corruption, and costs GPU time to find.  This model runs the
same loops as Python generators under a random scheduler that also delays the asynchronous completions arbitrarily, and checks

  * no deadlock: every agent runs to completion for every schedule tried;
  * no WAR hazard: a shared-memory slot is only refilled after every MMA that read it has retired;
  * no RAW hazard: an MMA only reads a slot whose bytes have landed and belong to the (tile, channel block, tap) it expects;
  * accumulators: the MMA only overwrites a TMEM stage the epilogue has drained, the epilogue only reads a completed stage of the
    tile it expects, and every tile is drained exactly once by exactly one owner.

Covered variants: the validated schedule (two epilogue groups splitting a tile's chunks), `kEpiAlt` (groups take alternate tiles),
weights-stationary modes (9 or 9 * kWS resident weight tiles), CTA pairs (both CTAs' producers feed the leader's barriers, multicast
commits), and the halo-load wgrad kernel's item loop.  It is a model of the PROTOCOL, written from the kernel source; it cannot prove
the CUDA code equals the model, but it did have to agree with the hardware-validated variants before the new ones were trusted.
"""
import random

import pytest


class Bar:
    """mbarrier: `count` arrivals (+ zero outstanding transaction bytes) complete a phase."""

    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count

    def arrive(self, n=1):
        assert self.pending >= n, "more arrivals than the barrier was initialised for"
        self.pending -= n
        self._maybe_flip()

    def expect_tx(self, nbytes):           # mbarrier.arrive.expect_tx
        self.tx += nbytes
        self.arrive(1)

    def complete_tx(self, nbytes):
        # the transaction count is signed in hardware: in CTA-pair kernels the peer's bytes may land on the leader's barrier before
        # the leader's own arrive.expect_tx; the phase still cannot complete early because that arrival is outstanding
        self.tx -= nbytes
        self._maybe_flip()

    def passed(self, parity):              # try_wait.parity: has the phase with this parity completed?
        return self.phase != parity


class Sim:
    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.agents = []                   # [name, generator, waiting-on (bar, parity) or None]
        self.async_unordered = []          # TMA completions: may fire in any order
        self.async_fifo = []               # tcgen05.commit arrivals: fire in issue order
        self.steps = 0

    def add(self, name, gen):
        self.agents.append([name, gen, None])

    def run(self, limit=2_000_000):
        live = list(self.agents)
        while live or self.async_unordered or self.async_fifo:
            self.steps += 1
            assert self.steps < limit, "model did not terminate"
            choices = []
            for a in live:
                if a[2] is None or a[2][0].passed(a[2][1]):
                    choices.append(("agent", a))
            if self.async_unordered:
                choices.append(("tma", None))
            if self.async_fifo:
                choices.append(("commit", None))
            if not choices:
                raise AssertionError("deadlock: " + ", ".join(f"{a[0]} waits" for a in live))
            kind, a = self.rng.choice(choices)
            if kind == "tma":
                self.async_unordered.pop(self.rng.randrange(len(self.async_unordered)))()
            elif kind == "commit":
                self.async_fifo.pop(0)()
            else:
                a[2] = None
                try:
                    w = next(a[1])
                    a[2] = w               # None = just yield the processor, (bar, parity) = wait
                except StopIteration:
                    live.remove(a)


def halo_model(seed, tiles, cblocks, chunks, alt, bstat, kb_stages, pair=False, ws=0):
    """conv3x3_halo_kernel: `tiles` tiles per worker, `chunks` = BN / 64 epilogue chunks per tile."""
    sim = Sim(seed)
    nct = 2 if pair else 1
    groups = 2
    active_groups = 2 if alt else min(chunks, 2)
    ka = 2
    a_full = [Bar(1) for _ in range(ka)]                                  # leader's
    b_full = [Bar(1) for _ in range(kb_stages)]                           # leader's
    a_empty = [[Bar(1) for _ in range(ka)] for _ in range(nct)]           # per CTA (multicast commit)
    b_empty = [[Bar(1) for _ in range(kb_stages)] for _ in range(nct)]
    tmem_full = [[Bar(1) for _ in range(2)] for _ in range(nct)]
    tmem_empty = [Bar(nct * 128 * min(chunks, 2)) for _ in range(2)]      # leader's: the kernel initialises it with kActiveGroups = min(chunks, 2)
    a_slot = [[None] * ka for _ in range(nct)]                            # content tag or None while in flight / free
    b_slot = [[None] * kb_stages for _ in range(nct)]
    a_busy = [[0] * ka for _ in range(nct)]                               # MMAs issued on the slot and not yet retired
    b_busy = [[0] * kb_stages for _ in range(nct)]
    acc_state = [{"tile": None, "complete": False} for _ in range(2)]
    drains = [None, None]                                                 # epilogue arrivals since the MMA last claimed the stage (None = never used)
    owners_per_stage = nct * (1 if alt else active_groups)
    drained = []
    nres = 9 * ws if ws else 9

    def commit(fn):
        sim.async_fifo.append(fn)

    def producer(cta):
        leader = cta == 0
        sa = pa = sb = pb = 0

        def load_b(slot, tag):
            assert b_busy[cta][slot] == 0, "WAR: weight slot refilled while an MMA still reads it"
            b_slot[cta][slot] = None
            sim.async_unordered.append(lambda: (b_slot[cta].__setitem__(slot, tag), b_full[slot].complete_tx(1)))

        if bstat and tiles:
            for s in range(nres):
                if leader:
                    b_full[s].expect_tx(nct)
                load_b(s, ("w", s))
                yield None
        for t in range(tiles):
            for cb in range(cblocks):
                yield (a_empty[cta][sa], pa ^ 1)
                if leader:
                    a_full[sa].expect_tx(nct)
                assert a_busy[cta][sa] == 0, "WAR: halo slot refilled while an MMA still reads it"
                a_slot[cta][sa] = None
                sim.async_unordered.append(lambda sa=sa, tag=(t, cb): (a_slot[cta].__setitem__(sa, tag), a_full[sa].complete_tx(1)))
                sa += 1
                if sa == ka:
                    sa, pa = 0, pa ^ 1
                if bstat:
                    continue
                for tap in range(9):
                    yield (b_empty[cta][sb], pb ^ 1)
                    if leader:
                        b_full[sb].expect_tx(nct)
                    load_b(sb, (t, cb, tap))
                    sb += 1
                    if sb == kb_stages:
                        sb, pb = 0, pb ^ 1

    def mma():
        sa = pa = sb = pb = 0
        acc = acc_phase = 0
        if bstat and tiles:
            for s in range(nres):
                yield (b_full[s], 0)
        for t in range(tiles):
            yield (tmem_empty[acc], acc_phase ^ 1)
            st = acc_state[acc]
            assert drains[acc] in (None, owners_per_stage), f"MMA overwrites an accumulator after {drains[acc]} of {owners_per_stage} drains"
            drains[acc] = 0
            st.update(tile=t, complete=False)
            for cb in range(cblocks):
                yield (a_full[sa], pa)
                for tap in range(9):
                    slot = (cb * 9 + tap if ws else tap) if bstat else sb
                    if not bstat:
                        yield (b_full[sb], pb)
                    for c in range(nct):
                        assert a_slot[c][sa] == (t, cb), f"RAW: halo slot holds {a_slot[c][sa]}, expected {(t, cb)}"
                        want = ("w", slot) if bstat else (t, cb, tap)
                        assert b_slot[c][slot] == want, f"RAW: weight slot holds {b_slot[c][slot]}, expected {want}"
                        a_busy[c][sa] += 1
                        b_busy[c][slot] += 1
                    commit(lambda sa=sa, slot=slot: [(a_busy[c].__setitem__(sa, a_busy[c][sa] - 1), b_busy[c].__setitem__(slot, b_busy[c][slot] - 1))
                                                    for c in range(nct)])      # the MMA retires
                    if not bstat:
                        commit(lambda sb=sb: [b_empty[c][sb].arrive() for c in range(nct)])
                        sb += 1
                        if sb == kb_stages:
                            sb, pb = 0, pb ^ 1
                    yield None
                commit(lambda sa=sa: [a_empty[c][sa].arrive() for c in range(nct)])
                sa += 1
                if sa == ka:
                    sa, pa = 0, pa ^ 1
            commit(lambda acc=acc: (acc_state[acc].__setitem__("complete", True), [tmem_full[c][acc].arrive() for c in range(nct)]))
            acc += 1
            if acc == 2:
                acc, acc_phase = 0, acc_phase ^ 1

    def epilogue(cta, grp):
        acc = acc_phase = 0
        if not (grp < chunks or alt):
            return
        for t in range(tiles):
            if alt and acc != grp:
                acc += 1
                if acc == 2:
                    acc, acc_phase = 0, acc_phase ^ 1
                continue
            yield (tmem_full[cta][acc], acc_phase)
            st = acc_state[acc]
            assert st["tile"] == t and st["complete"], f"epilogue reads accumulator of tile {st['tile']} (complete={st['complete']}), expected {t}"
            yield None                                              # TMEM -> registers
            drained.append((cta, grp, t))
            drains[acc] += 1
            tmem_empty[acc].arrive(128)
            acc += 1
            if acc == 2:
                acc, acc_phase = 0, acc_phase ^ 1

    for c in range(nct):
        sim.add(f"producer{c}", producer(c))
        for g in range(groups):
            sim.add(f"epilogue{c}.{g}", epilogue(c, g))
    sim.add("mma", mma())
    sim.run()
    # every tile drained by the expected owners, once each
    for c in range(nct):
        for t in range(tiles):
            owners = sorted(g for (cc, g, tt) in drained if cc == c and tt == t)
            expect = [t % 2] if alt else list(range(active_groups))
            assert owners == expect, (c, t, owners, expect)
    assert all(b == 0 for row in a_busy + b_busy for b in row)


def _cfgs():
    out = []
    for tiles in (0, 1, 2, 3, 5, 8):
        for cblocks in (1, 2):
            out.append(dict(tiles=tiles, cblocks=cblocks, chunks=2, alt=False, bstat=False, kb_stages=8))          # validated: BN = 128
            out.append(dict(tiles=tiles, cblocks=cblocks, chunks=4, alt=False, bstat=False, kb_stages=4))          # validated: BN = 256
            out.append(dict(tiles=tiles, cblocks=cblocks, chunks=1, alt=False, bstat=cblocks == 1, kb_stages=9))   # validated: BN = 64 (+ stationary weights)
            out.append(dict(tiles=tiles, cblocks=cblocks, chunks=1, alt=True, bstat=cblocks == 1, kb_stages=9))    # kEpiAlt
            out.append(dict(tiles=tiles, cblocks=cblocks, chunks=2, alt=False, bstat=False, kb_stages=9, pair=True))   # validated: CTA pair
        out.append(dict(tiles=tiles, cblocks=2, chunks=2, alt=False, bstat=True, kb_stages=18, pair=True, ws=2))   # kWS = 2 pair
        out.append(dict(tiles=tiles, cblocks=1, chunks=1, alt=True, bstat=True, kb_stages=9, pair=True))           # 64-column pair + kEpiAlt
        out.append(dict(tiles=tiles, cblocks=1, chunks=1, alt=False, bstat=True, kb_stages=9, pair=True))          # 64-column pair
    return out


@pytest.mark.parametrize("cfg", _cfgs(), ids=lambda c: "-".join(f"{k}{int(v) if isinstance(v, bool) else v}" for k, v in c.items()))
def test_halo_kernel_protocol(cfg):
    for seed in range(20):
        halo_model(seed, **cfg)


def wgrad_model(seed, items, tiles_per_item, stages=4):
    """wgrad3x3_halo_kernel: `items` work items per CTA, each accumulating `tiles_per_item` M tiles into one accumulator set."""
    sim = Sim(seed)
    full = [Bar(1) for _ in range(stages)]
    empty = [Bar(1) for _ in range(stages)]
    tmem_full, tmem_empty = Bar(1), Bar(128)
    slot = [None] * stages
    busy = [0] * stages
    acc = {"item": None, "tiles": 0, "complete": False, "drains": None}
    out = []

    def producer():
        s = ph = 0
        for it in range(items):
            for t in range(tiles_per_item):
                yield (empty[s], ph ^ 1)
                full[s].expect_tx(2)                                  # X halo box + dY box on one barrier
                assert busy[s] == 0, "WAR: stage refilled while MMAs still read it"
                slot[s] = None
                got = []
                for part in ("x", "dy"):
                    sim.async_unordered.append(lambda s=s, tag=(it, t), part=part, got=got:
                                               (got.append(part), slot.__setitem__(s, tag) if len(got) == 2 else None, full[s].complete_tx(1)))
                s += 1
                if s == stages:
                    s, ph = 0, ph ^ 1

    def mma():
        s = ph = item_phase = 0
        for it in range(items):
            yield (tmem_empty, item_phase ^ 1)
            assert acc["drains"] in (None, 1), "MMA overwrites accumulators the epilogue has not drained"
            acc.update(item=it, tiles=0, complete=False, drains=0)
            for t in range(tiles_per_item):
                yield (full[s], ph)
                assert slot[s] == (it, t), f"RAW: stage holds {slot[s]}, expected {(it, t)}"
                busy[s] += 1
                acc["tiles"] += 1
                sim.async_fifo.append(lambda s=s: (busy.__setitem__(s, busy[s] - 1), empty[s].arrive()))
                s += 1
                if s == stages:
                    s, ph = 0, ph ^ 1
                yield None
            sim.async_fifo.append(lambda: (acc.__setitem__("complete", True), tmem_full.arrive()))
            item_phase ^= 1

    def epilogue():
        item_phase = 0
        for it in range(items):
            yield (tmem_full, item_phase)
            assert acc["item"] == it and acc["complete"] and acc["tiles"] == tiles_per_item
            yield None
            out.append(it)
            acc["drains"] += 1
            tmem_empty.arrive(128)
            item_phase ^= 1

    sim.add("producer", producer())
    sim.add("mma", mma())
    sim.add("epilogue", epilogue())
    sim.run()
    assert out == list(range(items)) and all(b == 0 for b in busy)


@pytest.mark.parametrize("items,tiles", [(1, 1), (1, 7), (2, 3), (3, 5), (1, 48), (4, 1)])
def test_halo_wgrad_protocol(items, tiles):
    for seed in range(20):
        wgrad_model(seed, items, tiles)


def test_the_model_catches_a_wrong_arrive_count():
    """Sanity of the checker itself: with kEpiAlt but the barrier initialised for two groups' arrivals the pipeline must deadlock."""
    import types
    src = halo_model.__code__
    assert src is not None
    orig = Bar.__init__

    def bad_init(self, count):                       # every 128-thread barrier expects twice the arrivals it will get
        orig(self, count * 2 if count == 128 else count)

    Bar.__init__ = bad_init
    try:
        with pytest.raises(AssertionError, match="deadlock"):
            halo_model(0, tiles=3, cblocks=1, chunks=1, alt=True, bstat=True, kb_stages=9)
    finally:
        Bar.__init__ = orig
        del types
