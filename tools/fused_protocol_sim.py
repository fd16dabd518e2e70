"""Protocol model of conv_fused_mid_kernel (rusty_sr_b200/csrc/conv_fused.cuh), runnable without a GPU.

Every warp role of the kernel is a generator that mirrors the kernel's loops, ring indices and mbarrier
parities line by line; mbarriers (phase bit + arrival count), the in-order tcgen05 pipe (MMAs and commits
retire in issue order, asynchronously to the issuing thread) and the TMA engine (loads complete after a
random delay) are modelled with a randomised scheduler.  Shared-memory slots, TMEM slots and tiles carry
tags saying what they hold, and every read asserts it sees the data the algorithm expects — so a wrong
slot index, a wrong parity, a lapped barrier, a missing hand-off or a deadlock shows up here instead of
as a stuck kernel on the GPU.  tests/test_fused_protocol.py runs it over many item shapes and seeds.
"""
from __future__ import annotations

import random
from collections import deque

NR1, NR2, NB, NT, LOOKAHEAD = 3, 3, 4, 5, 4   # single-CTA kernel: FusedCfgT<false>


class Bar:
    def __init__(self, count, name):
        self.count, self.pending, self.done, self.name = count, count, 0, name

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"over-arrival on {self.name}"
        if self.pending == 0:
            self.pending = self.count
            self.done += 1

    def test(self, parity):      # mbarrier.try_wait.parity: has the phase with this parity completed?
        return (self.done & 1) != parity


class Deadlock(Exception):
    pass


class Sim:
    def __init__(self, items, seed=0, queue_depth=8):
        self.rng = random.Random(seed)
        self.items = items                      # list of n (rows per work item)
        B = lambda c, n: Bar(c, n)
        self.a_full = [B(1, f"a_full{i}") for i in range(NR1)]
        self.a_empty = [B(1, f"a_empty{i}") for i in range(NR1)]
        self.b_full = [B(1, f"b_full{i}") for i in range(NB)]
        self.b_empty = [B(1, f"b_empty{i}") for i in range(NB)]
        self.p2_full, self.p2_empty = B(1, "p2_full"), B(4, "p2_empty")
        self.p3_full, self.p3_empty = B(1, "p3_full"), B(4, "p3_empty")
        self.t1_full = [B(1, f"t1_full{i}") for i in range(NT)]
        self.t2_full = [B(1, f"t2_full{i}") for i in range(NT)]
        self.t_empty = [B(4, f"t_empty{i}") for i in range(NT)]
        self.l2_full = [B(4, f"l2_full{i}") for i in range(NR2)]
        self.l2_empty = [B(1, f"l2_empty{i}") for i in range(NR2)]
        # storage tags
        self.sA = [None] * NR1                   # (item, idx)
        self.sL = [[None] * 4 for _ in range(NR2)]   # per E1 warp quarter: (item, q)
        self.sB = [None] * NB                    # (kind, item, op, tap)
        self.sP2, self.sP3 = None, None
        self.T = [dict(l2pre=None, s1=None, s1_taps=0, s2=None, s2_taps=0) for _ in range(NT)]
        self.pipe = deque()                      # in-order tensor pipe: ('mma', fn) | ('commit', bar)
        self.queue_depth = queue_depth
        self.tma = []                            # pending loads: [delay, fn]
        self.outputs = []                        # (item, r) rows stored by E2 warp 0
        self.l2_written = []

    # ---- agents (generators yield a predicate to wait for, or None to just reschedule) ----
    def wait(self, bar, parity):
        return lambda: bar.test(parity)

    def tma_load(self, fn, bar):
        def done():
            fn()
            bar.arrive()
        self.tma.append([self.rng.randint(1, 12), done])

    def l1_producer(self):
        a = 0
        for it, n in enumerate(self.items):
            for i in range(n + 4):
                sl, par = a % NR1, ((a // NR1) & 1) ^ 1
                yield self.wait(self.a_empty[sl], par)
                self.tma_load(lambda sl=sl, it=it, i=i: self.sA.__setitem__(sl, (it, i)), self.a_full[sl])
                a += 1

    def ops(self, n):
        """The MMA warp's op order (must match the weight producer)."""
        ns1 = n + 2
        i1 = r = 0
        while r < n:
            if i1 < ns1 and i1 < r + LOOKAHEAD:
                yield ("s1", i1)
                i1 += 1
            else:
                yield ("s2", r)
                r += 1
        assert i1 == ns1, "not every S1 row was issued"

    def weight_producer(self):
        st, par = 0, 1
        for it, n in enumerate(self.items):
            for kind, idx in self.ops(n):
                for t in range(9):
                    yield self.wait(self.b_empty[st], par)
                    self.tma_load(lambda st=st, v=(kind, it, idx, t): self.sB.__setitem__(st, v), self.b_full[st])
                    st += 1
                    if st == NB:
                        st, par = 0, par ^ 1

    def p2_producer(self):
        u = 0
        for it, n in enumerate(self.items):
            for q in range(n + 2):
                yield self.wait(self.p2_empty, (u & 1) ^ 1)
                self.tma_load(lambda v=(it, q): setattr(self, "sP2", v), self.p2_full)
                u += 1

    def p3_producer(self):
        u = 0
        for it, n in enumerate(self.items):
            for r in range(n):
                yield self.wait(self.p3_empty, (u & 1) ^ 1)
                self.tma_load(lambda v=(it, r): setattr(self, "sP3", v), self.p3_full)
                u += 1

    def issue(self, entry):
        """Issuing thread blocks while the MMA queue is full."""
        while len(self.pipe) >= self.queue_depth:
            yield lambda: len(self.pipe) < self.queue_depth
        self.pipe.append(entry)

    def mma_issuer(self):
        c1_item = a_item = a_wait = l2_wait = c2 = 0
        bst, bpar = 0, 0
        for it, n in enumerate(self.items):
            for kind, idx in self.ops(n):
                if kind == "s1":
                    q = idx
                    c = c1_item + q
                    ts = c % NT
                    yield self.wait(self.t_empty[ts], ((c // NT) & 1) ^ 1)
                    while a_wait <= a_item + q + 2:
                        yield self.wait(self.a_full[a_wait % NR1], (a_wait // NR1) & 1)
                        a_wait += 1
                    arow = [(a_item + q + ta) % NR1 for ta in range(3)]
                    for t in range(9):
                        ta = t // 3
                        yield self.wait(self.b_full[bst], bpar)

                        def mma(ts=ts, sl=arow[ta], bs=bst, it=it, q=q, ta=ta, t=t):
                            assert self.sA[sl] == (it, q + ta), f"S1({it},{q}) tap row {ta}: l1 slot {sl} holds {self.sA[sl]}"
                            assert self.sB[bs] == ("s1", it, q, t), f"S1({it},{q}) tap {t}: weight stage holds {self.sB[bs]}"
                            T = self.T[ts]
                            if t == 0:
                                assert T["s1"] is None, f"S1({it},{q}) overwrites live T slot {ts}: {T}"
                                T.update(l2pre=(it, q), s1=(it, q), s1_taps=0, s2=None, s2_taps=0)
                            assert T["s1"] == (it, q)
                            T["s1_taps"] += 1
                        yield from self.issue(("mma", mma))
                        if t == 2:
                            yield from self.issue(("commit", self.a_empty[(a_item + q) % NR1]))
                        yield from self.issue(("commit", self.b_empty[bst]))
                        bst += 1
                        if bst == NB:
                            bst, bpar = 0, bpar ^ 1
                    yield from self.issue(("commit", self.t1_full[ts]))
                    if q == n + 1:
                        yield from self.issue(("commit", self.a_empty[(a_item + q + 1) % NR1]))
                        yield from self.issue(("commit", self.a_empty[(a_item + q + 2) % NR1]))
                else:
                    r = idx
                    ts = (c1_item + r + 1) % NT
                    while l2_wait <= c1_item + r + 2:
                        yield self.wait(self.l2_full[l2_wait % NR2], (l2_wait // NR2) & 1)
                        l2_wait += 1
                    lrow = [(c1_item + r + ta) % NR2 for ta in range(3)]
                    for t in range(9):
                        ta = t // 3
                        yield self.wait(self.b_full[bst], bpar)

                        def mma(ts=ts, sl=lrow[ta], bs=bst, it=it, r=r, ta=ta, t=t):
                            assert self.sL[sl] == [(it, r + ta)] * 4, f"S2({it},{r}) tap row {ta}: l2 slot {sl} holds {self.sL[sl]}"
                            assert self.sB[bs] == ("s2", it, r, t), f"S2({it},{r}) tap {t}: weight stage holds {self.sB[bs]}"
                            T = self.T[ts]
                            assert T["s1"] == (it, r + 1) and T["s1_taps"] == 9, f"S2({it},{r}) onto T slot {ts}: {T}"
                            if t == 0:
                                assert T["s2"] is None
                                T["s2"] = (it, r)
                            T["s2_taps"] += 1
                        yield from self.issue(("mma", mma))
                        if t == 2:
                            yield from self.issue(("commit", self.l2_empty[(c1_item + r) % NR2]))
                        yield from self.issue(("commit", self.b_empty[bst]))
                        bst += 1
                        if bst == NB:
                            bst, bpar = 0, bpar ^ 1
                    yield from self.issue(("commit", self.t2_full[c2 % NT]))
                    if r == n - 1:
                        yield from self.issue(("commit", self.l2_empty[(c1_item + r + 1) % NR2]))
                        yield from self.issue(("commit", self.l2_empty[(c1_item + r + 2) % NR2]))
                    c2 += 1
            c1_item += n + 2
            a_item += n + 4

    def e1(self, ew):
        c = 0
        for it, n in enumerate(self.items):
            for q in range(n + 2):
                ts, ls = c % NT, c % NR2
                yield self.wait(self.t1_full[ts], (c // NT) & 1)
                T = self.T[ts]
                assert T["l2pre"] == (it, q) and T["s1_taps"] == 9, f"E1({it},{q}) reads T slot {ts}: {T}"
                yield self.wait(self.p2_full, c & 1)
                assert self.sP2 == (it, q), f"E1({it},{q}) P2 tile holds {self.sP2}"
                yield None
                self.p2_empty.arrive()
                yield self.wait(self.l2_empty[ls], ((c // NR2) & 1) ^ 1)
                self.sL[ls][ew] = (it, q)
                if ew == 0:
                    self.l2_written.append((it, q))
                yield None
                self.l2_full[ls].arrive()
                if q == 0 or q == n + 1:
                    if ew == 0:
                        T["s1"] = None           # slot released (model: first releasing warp clears the tag)
                    self.t_empty[ts].arrive()
                c += 1

    def e2(self, ew):
        c1_item = c2 = 0
        for it, n in enumerate(self.items):
            for r in range(n):
                ts = (c1_item + r + 1) % NT
                yield self.wait(self.t2_full[c2 % NT], (c2 // NT) & 1)
                T = self.T[ts]
                assert T["s1"] == (it, r + 1) and T["s1_taps"] == 9 and T["s2"] == (it, r) and T["s2_taps"] == 9, \
                    f"E2({it},{r}) reads T slot {ts}: {T}"
                yield self.wait(self.p3_full, c2 & 1)
                assert self.sP3 == (it, r), f"E2({it},{r}) P3 tile holds {self.sP3}"
                yield None
                self.p3_empty.arrive()
                yield None
                if ew == 0:
                    self.outputs.append((it, r))
                # all four warps must have read before the slot is reused: t_empty has count 4
                self.t_empty[ts].arrive()
                if self.t_empty[ts].pending == self.t_empty[ts].count:   # this arrival completed the phase
                    T["s1"] = None
                c2 += 1
            c1_item += n + 2

    # ---- scheduler ----
    def run(self, max_steps=10_000_000):
        agents = {"l1": self.l1_producer(), "w": self.weight_producer(), "p2": self.p2_producer(),
                  "p3": self.p3_producer(), "mma": self.mma_issuer()}
        for i in range(4):
            agents[f"e1_{i}"] = self.e1(i)
            agents[f"e2_{i}"] = self.e2(i)
        blocked = {k: None for k in agents}     # predicate each agent waits on
        live = set(agents)
        steps = 0
        while live or self.pipe or self.tma:
            steps += 1
            assert steps < max_steps
            choices = [k for k in live if blocked[k] is None or blocked[k]()]
            if self.pipe:
                choices.append("#pipe")
            if self.tma:
                choices.append("#tma")
            if not choices:
                raise Deadlock({k: getattr(blocked[k], "__closure__", None) and
                                [getattr(c.cell_contents, "name", c.cell_contents) for c in blocked[k].__closure__]
                                for k in live})
            k = self.rng.choice(choices)
            if k == "#pipe":
                kind, x = self.pipe.popleft()
                x() if kind == "mma" else x.arrive()
            elif k == "#tma":
                j = self.rng.randrange(len(self.tma))
                self.tma[j][0] -= 1
                if self.tma[j][0] <= 0:
                    self.tma.pop(j)[1]()
            else:
                blocked[k] = None
                try:
                    blocked[k] = next(agents[k])
                except StopIteration:
                    live.discard(k)
        want = [(it, r) for it, n in enumerate(self.items) for r in range(n)]
        assert self.outputs == want, "E2 did not store every row exactly once, in order"
        want_l2 = [(it, q) for it, n in enumerate(self.items) for q in range(n + 2)]
        assert self.l2_written == want_l2
        return steps


# ---------------------------------------------------------------------------------------------------------------
# CTA-pair variant (template parameter PAIR of conv_fused_mid_kernel; the library's default plan).  Two CTAs, each with
# its own rings, tiles, tensor memory, producers and epilogue warps, walk the same number of rows in lock step; ONE MMA
# issuer (the leader's) multiplies both CTAs' rows at once.  Hand-offs that cross the pair:
#   * l1 rows / weight halves: each CTA's TMA signals the LEADER's a_full / b_full (modelled: one arrival per CTA);
#   * tcgen05.commit ... multicast::cluster: arrives on the same barrier of BOTH CTAs (modelled as two deliveries that
#     land independently, in either order, some time after the commit retires);
#   * the epilogue warps of both CTAs arrive on the leader's l2_full / t_empty (count 8).
# Weight stages hold TPS taps (PAIR_TPS) in a ring of PAIR_NB stages; the l1 ring has PAIR_NR1 slots.
PAIR_TPS, PAIR_NB, PAIR_NR1 = 3, 3, 3


class PairSim:
    def __init__(self, items, seed=0, queue_depth=8):
        self.rng = random.Random(seed)
        self.items = items                      # rows per CTA and work item (both CTAs run the same count)
        self.NSTG = 9 // PAIR_TPS
        B = lambda c, n: Bar(c, n)
        two = lambda f: [f(0), f(1)]
        # leader-only barriers
        self.a_full = [B(2, f"a_full{i}") for i in range(PAIR_NR1)]
        self.b_full = [B(2, f"b_full{i}") for i in range(PAIR_NB)]
        self.l2_full = [B(8, f"l2_full{i}") for i in range(NR2)]
        self.t_empty = [B(8, f"t_empty{i}") for i in range(NT)]
        # per-CTA barriers
        self.a_empty = two(lambda c: [B(1, f"a_empty{i}@{c}") for i in range(PAIR_NR1)])
        self.b_empty = two(lambda c: [B(1, f"b_empty{i}@{c}") for i in range(PAIR_NB)])
        self.t1_full = two(lambda c: [B(1, f"t1_full{i}@{c}") for i in range(NT)])
        self.t2_full = two(lambda c: [B(1, f"t2_full{i}@{c}") for i in range(NT)])
        self.l2_empty = two(lambda c: [B(1, f"l2_empty{i}@{c}") for i in range(NR2)])
        self.p2_full, self.p2_empty = two(lambda c: B(1, f"p2_full@{c}")), two(lambda c: B(4, f"p2_empty@{c}"))
        self.p3_full, self.p3_empty = two(lambda c: B(1, f"p3_full@{c}")), two(lambda c: B(4, f"p3_empty@{c}"))
        # per-CTA storage tags
        self.sA = two(lambda c: [None] * PAIR_NR1)
        self.sL = two(lambda c: [[None] * 4 for _ in range(NR2)])
        self.sB = two(lambda c: [None] * PAIR_NB)
        self.sP2, self.sP3 = [None, None], [None, None]
        self.T = two(lambda c: [dict(l2pre=None, s1=None, s1_taps=0, s2=None, s2_taps=0) for _ in range(NT)])
        self.pipe = deque()
        self.queue_depth = queue_depth
        self.tma = []                            # pending asynchronous deliveries: [delay, fn]
        self.outputs = [[], []]
        self.l2_written = [[], []]

    def wait(self, bar, parity):
        return lambda: bar.test(parity)

    def later(self, fn, lo=1, hi=12):
        self.tma.append([self.rng.randint(lo, hi), fn])

    def commit(self, bars):
        """tcgen05.commit multicast: one pipe entry; when it retires, each CTA's barrier gets its arrival on its own."""
        return ("commit", bars)

    # ---- producers (run in both CTAs) ----
    def l1_producer(self, cta):
        a = 0
        for it, n in enumerate(self.items):
            for i in range(n + 4):
                sl, par = a % PAIR_NR1, ((a // PAIR_NR1) & 1) ^ 1
                yield self.wait(self.a_empty[cta][sl], par)

                def done(sl=sl, it=it, i=i):
                    self.sA[cta][sl] = (it, i)
                    self.a_full[sl].arrive()         # complete_tx on the LEADER's barrier
                self.later(done)
                a += 1

    def ops(self, n):
        ns1 = n + 2
        i1 = r = 0
        while r < n:
            if i1 < ns1 and i1 < r + LOOKAHEAD:
                yield ("s1", i1)
                i1 += 1
            else:
                yield ("s2", r)
                r += 1

    def weight_producer(self, cta):
        st, par = 0, 1
        for it, n in enumerate(self.items):
            for kind, idx in self.ops(n):
                for sp in range(self.NSTG):
                    yield self.wait(self.b_empty[cta][st], par)

                    def done(st=st, v=(kind, it, idx, sp)):
                        self.sB[cta][st] = v
                        self.b_full[st].arrive()
                    self.later(done)
                    st += 1
                    if st == PAIR_NB:
                        st, par = 0, par ^ 1

    def p2_producer(self, cta):
        u = 0
        for it, n in enumerate(self.items):
            for q in range(n + 2):
                yield self.wait(self.p2_empty[cta], (u & 1) ^ 1)

                def done(v=(it, q)):
                    self.sP2[cta] = v
                    self.p2_full[cta].arrive()
                self.later(done)
                u += 1

    def p3_producer(self, cta):
        u = 0
        for it, n in enumerate(self.items):
            for r in range(n):
                yield self.wait(self.p3_empty[cta], (u & 1) ^ 1)

                def done(v=(it, r)):
                    self.sP3[cta] = v
                    self.p3_full[cta].arrive()
                self.later(done)
                u += 1

    def issue(self, entry):
        while len(self.pipe) >= self.queue_depth:
            yield lambda: len(self.pipe) < self.queue_depth
        self.pipe.append(entry)

    # ---- the leader's MMA issuer ----
    def mma_issuer(self):
        both = lambda bars, i: [bars[0][i], bars[1][i]]
        c1_item = a_item = a_wait = l2_wait = c2 = 0
        bst, bpar = 0, 0
        for it, n in enumerate(self.items):
            for kind, idx in self.ops(n):
                if kind == "s1":
                    q = idx
                    c = c1_item + q
                    ts = c % NT
                    yield self.wait(self.t_empty[ts], ((c // NT) & 1) ^ 1)
                    while a_wait <= a_item + q + 2:
                        yield self.wait(self.a_full[a_wait % PAIR_NR1], (a_wait // PAIR_NR1) & 1)
                        a_wait += 1
                    arow = [(a_item + q + ta) % PAIR_NR1 for ta in range(3)]
                    for sp in range(self.NSTG):
                        yield self.wait(self.b_full[bst], bpar)
                        for tt in range(PAIR_TPS):
                            t = sp * PAIR_TPS + tt
                            ta = t // 3

                            def mma(ts=ts, sl=arow[ta], bs=bst, it=it, q=q, ta=ta, t=t, sp=sp):
                                for cta in (0, 1):   # one cta_group::2 MMA reads the same slots of both CTAs
                                    assert self.sA[cta][sl] == (it, q + ta), f"S1({it},{q}) tap row {ta} cta {cta}: l1 slot {sl} holds {self.sA[cta][sl]}"
                                    assert self.sB[cta][bs] == ("s1", it, q, sp), f"S1({it},{q}) stage {sp} cta {cta}: weights hold {self.sB[cta][bs]}"
                                    T = self.T[cta][ts]
                                    if t == 0:
                                        assert T["s1"] is None, f"S1({it},{q}) overwrites live T slot {ts} of cta {cta}: {T}"
                                        T.update(l2pre=(it, q), s1=(it, q), s1_taps=0, s2=None, s2_taps=0)
                                    assert T["s1"] == (it, q)
                                    T["s1_taps"] += 1
                            yield from self.issue(("mma", mma))
                            if t == 2:
                                yield from self.issue(self.commit(both(self.a_empty, (a_item + q) % PAIR_NR1)))
                        yield from self.issue(self.commit(both(self.b_empty, bst)))
                        bst += 1
                        if bst == PAIR_NB:
                            bst, bpar = 0, bpar ^ 1
                    yield from self.issue(self.commit(both(self.t1_full, ts)))
                    if q == n + 1:
                        yield from self.issue(self.commit(both(self.a_empty, (a_item + q + 1) % PAIR_NR1)))
                        yield from self.issue(self.commit(both(self.a_empty, (a_item + q + 2) % PAIR_NR1)))
                else:
                    r = idx
                    ts = (c1_item + r + 1) % NT
                    while l2_wait <= c1_item + r + 2:
                        yield self.wait(self.l2_full[l2_wait % NR2], (l2_wait // NR2) & 1)
                        l2_wait += 1
                    lrow = [(c1_item + r + ta) % NR2 for ta in range(3)]
                    for sp in range(self.NSTG):
                        yield self.wait(self.b_full[bst], bpar)
                        for tt in range(PAIR_TPS):
                            t = sp * PAIR_TPS + tt
                            ta = t // 3

                            def mma(ts=ts, sl=lrow[ta], bs=bst, it=it, r=r, ta=ta, t=t, sp=sp):
                                for cta in (0, 1):
                                    assert self.sL[cta][sl] == [(it, r + ta)] * 4, f"S2({it},{r}) tap row {ta} cta {cta}: l2 slot {sl} holds {self.sL[cta][sl]}"
                                    assert self.sB[cta][bs] == ("s2", it, r, sp), f"S2({it},{r}) stage {sp} cta {cta}: weights hold {self.sB[cta][bs]}"
                                    T = self.T[cta][ts]
                                    assert T["s1"] == (it, r + 1) and T["s1_taps"] == 9, f"S2({it},{r}) onto T slot {ts} of cta {cta}: {T}"
                                    if t == 0:
                                        assert T["s2"] is None
                                        T["s2"] = (it, r)
                                    T["s2_taps"] += 1
                            yield from self.issue(("mma", mma))
                            if t == 2:
                                yield from self.issue(self.commit(both(self.l2_empty, (c1_item + r) % NR2)))
                        yield from self.issue(self.commit(both(self.b_empty, bst)))
                        bst += 1
                        if bst == PAIR_NB:
                            bst, bpar = 0, bpar ^ 1
                    yield from self.issue(self.commit(both(self.t2_full, c2 % NT)))
                    if r == n - 1:
                        yield from self.issue(self.commit(both(self.l2_empty, (c1_item + r + 1) % NR2)))
                        yield from self.issue(self.commit(both(self.l2_empty, (c1_item + r + 2) % NR2)))
                    c2 += 1
            c1_item += n + 2
            a_item += n + 4

    # ---- epilogue warps (both CTAs) ----
    def e1(self, cta, ew):
        c = 0
        for it, n in enumerate(self.items):
            for q in range(n + 2):
                ts, ls = c % NT, c % NR2
                yield self.wait(self.t1_full[cta][ts], (c // NT) & 1)
                T = self.T[cta][ts]
                assert T["l2pre"] == (it, q) and T["s1_taps"] == 9, f"E1({it},{q}) cta {cta} reads T slot {ts}: {T}"
                yield self.wait(self.p2_full[cta], c & 1)
                assert self.sP2[cta] == (it, q), f"E1({it},{q}) cta {cta} P2 tile holds {self.sP2[cta]}"
                yield None
                self.p2_empty[cta].arrive()
                yield self.wait(self.l2_empty[cta][ls], ((c // NR2) & 1) ^ 1)
                self.sL[cta][ls][ew] = (it, q)
                if ew == 0:
                    self.l2_written[cta].append((it, q))
                yield None
                self.later(self.l2_full[ls].arrive, 1, 4)        # (remote) arrive on the leader's barrier
                if q == 0 or q == n + 1:
                    if ew == 0:
                        T["s1"] = None
                    self.later(self.t_empty[ts].arrive, 1, 4)
                c += 1

    def e2(self, cta, ew):
        c1_item = c2 = 0
        for it, n in enumerate(self.items):
            for r in range(n):
                ts = (c1_item + r + 1) % NT
                yield self.wait(self.t2_full[cta][c2 % NT], (c2 // NT) & 1)
                T = self.T[cta][ts]
                assert T["s1"] == (it, r + 1) and T["s1_taps"] == 9 and T["s2"] == (it, r) and T["s2_taps"] == 9, \
                    f"E2({it},{r}) cta {cta} reads T slot {ts}: {T}"
                yield self.wait(self.p3_full[cta], c2 & 1)
                assert self.sP3[cta] == (it, r), f"E2({it},{r}) cta {cta} P3 tile holds {self.sP3[cta]}"
                yield None
                self.p3_empty[cta].arrive()
                yield None
                if ew == 0:
                    self.outputs[cta].append((it, r))
                T["readers"] = T.get("readers", 0) + 1
                if T["readers"] == 4:                # all four warps of this CTA are done with the slot
                    T["readers"] = 0
                    T["s1"] = None
                self.later(self.t_empty[ts].arrive, 1, 4)
                c2 += 1
            c1_item += n + 2

    def run(self, max_steps=20_000_000):
        agents = {"mma": self.mma_issuer()}
        for cta in (0, 1):
            agents[f"l1@{cta}"] = self.l1_producer(cta)
            agents[f"w@{cta}"] = self.weight_producer(cta)
            agents[f"p2@{cta}"] = self.p2_producer(cta)
            agents[f"p3@{cta}"] = self.p3_producer(cta)
            for i in range(4):
                agents[f"e1_{i}@{cta}"] = self.e1(cta, i)
                agents[f"e2_{i}@{cta}"] = self.e2(cta, i)
        blocked = {k: None for k in agents}
        live = set(agents)
        steps = 0
        while live or self.pipe or self.tma:
            steps += 1
            assert steps < max_steps
            choices = [k for k in sorted(live) if blocked[k] is None or blocked[k]()]
            if self.pipe:
                choices.append("#pipe")
            if self.tma:
                choices.append("#tma")
            if not choices:
                raise Deadlock(sorted(live))
            k = self.rng.choice(choices)
            if k == "#pipe":
                kind, x = self.pipe.popleft()
                if kind == "mma":
                    x()
                else:
                    for bar in x:                    # multicast: each CTA's arrival lands on its own
                        self.later(bar.arrive, 1, 6)
            elif k == "#tma":
                j = self.rng.randrange(len(self.tma))
                self.tma[j][0] -= 1
                if self.tma[j][0] <= 0:
                    self.tma.pop(j)[1]()
            else:
                blocked[k] = None
                try:
                    blocked[k] = next(agents[k])
                except StopIteration:
                    live.discard(k)
        want = [(it, r) for it, n in enumerate(self.items) for r in range(n)]
        want_l2 = [(it, q) for it, n in enumerate(self.items) for q in range(n + 2)]
        for cta in (0, 1):
            assert self.outputs[cta] == want, f"cta {cta}: E2 did not store every row exactly once, in order"
            assert self.l2_written[cta] == want_l2
        return steps


if __name__ == "__main__":
    import sys
    n_runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = random.Random(1)
    for k in range(n_runs):
        items = [rng.choice([1, 1, 2, 3, 4, 5, 6, 7, 11, 23]) for _ in range(rng.randint(1, 3))]
        Sim(items, seed=k, queue_depth=rng.choice([1, 2, 8])).run()
    print(f"fused protocol model: {n_runs} randomised runs ok")
    for k in range(n_runs // 4):
        items = [rng.choice([1, 1, 2, 3, 4, 5, 6, 7, 11, 23]) for _ in range(rng.randint(1, 3))]
        PairSim(items, seed=k, queue_depth=rng.choice([1, 2, 8])).run()
    print(f"fused protocol model, CTA pairs: {n_runs // 4} randomised runs ok")
