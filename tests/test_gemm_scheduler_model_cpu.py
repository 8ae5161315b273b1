"""Executable model of the work-item hand-over inside gemm_tc_kernel's dynamic schedule (csrc/gemm_tc.cu): the leader's producer
warp draws items from a global counter (two fetches ahead), publishes item i + 1 when it starts item i through a 16-entry
ring, never publishes past the terminating item; the MMA warp, the epilogue and the epilogue's operand-box prefetch (which
looks IN_SLOTS items ahead) read the ring.  The roles are coroutines driven by a random scheduler under the kernel's real
back-pressure (operand ring of STAGES k-blocks, two accumulator stages).  Invariants checked for many shapes and seeds:
every role sees the same item sequence, each item exactly once across CTAs, no ring entry is overwritten before its readers
are done with it, nobody waits for an entry that is never published, and everything terminates."""
import random

RING = 16


class Cta:
    def __init__(self, sim, stages, kblocks, in_slots):
        self.sim, self.stages, self.kb, self.in_slots = sim, stages, kblocks, in_slots
        self.slot_index = [-1] * RING          # index last published into each ring entry
        self.slot_value = [None] * RING
        self.loaded = 0                        # k-blocks the producer has issued
        self.consumed = 0                      # k-blocks the MMA warp has consumed
        self.tiles_mma = 0                     # accumulators completed
        self.tiles_epi = 0                     # accumulators drained
        self.seen = {"producer": [], "mma": [], "epilogue": [], "prefetch": []}

    def publish(self, i, v):
        s = i % RING
        assert self.slot_index[s] in (-1, i - RING), "ring entry published out of order"
        # every reader of the previous occupant must be past it
        prev = i - RING
        if prev >= 0:
            for role, pos in self.sim.positions(self).items():
                assert pos > prev, f"ring overrun: {role} still needs item {prev} when {i} is published"
        self.slot_index[s], self.slot_value[s] = i, v

    def ready(self, i):
        return self.slot_index[i % RING] >= i

    def read(self, i):
        assert self.slot_index[i % RING] == i, "stale / overwritten ring entry"
        return self.slot_value[i % RING]


class Sim:
    def __init__(self, total, n_ctas, stages, kblocks, in_slots, rng):
        self.total, self.counter, self.rng = total, 0, rng
        self.ctas = [Cta(self, stages, kblocks, in_slots) for _ in range(n_ctas)]
        self.pos = {}

    def positions(self, cta):
        return {r: p for (c, r), p in self.pos.items() if c is cta}

    def fetch(self):
        v = self.counter
        self.counter += 1
        return v

    # ---- roles (generators: `yield` = blocked or pre-empted)
    def producer(self, c):
        t_cur, t_nxt, t_fly = self.fetch(), self.fetch(), self.fetch()
        c.publish(0, t_cur)
        it = 0
        while True:
            self.pos[(c, "producer")] = it
            if t_cur < self.total:
                c.publish(it + 1, t_nxt)
            if t_cur >= self.total:
                break
            c.seen["producer"].append(t_cur)
            for _ in range(c.kb):
                while c.loaded - c.consumed >= c.stages:     # `empty` barrier of the operand ring
                    yield
                c.loaded += 1
                yield
            t_cur, t_nxt, t_fly = t_nxt, t_fly, self.fetch()
            it += 1
        self.pos[(c, "producer")] = 1 << 30

    def mma(self, c):
        it = 0
        while True:
            self.pos[(c, "mma")] = it
            while not c.ready(it):
                yield
            t = c.read(it)
            if t >= self.total:
                break
            c.seen["mma"].append(t)
            while c.tiles_mma - c.tiles_epi >= 2:            # two accumulator stages
                yield
            for _ in range(c.kb):
                while c.consumed >= c.loaded:                # `full` barrier
                    yield
                c.consumed += 1
                yield
            c.tiles_mma += 1
            it += 1
        self.pos[(c, "mma")] = 1 << 30

    def epilogue(self, c):
        ended = [False]
        nxt = [0]

        def issue_item(g):                                   # the group's issuer thread: sequential g, stops at the terminator
            assert g == nxt[0]
            nxt[0] += 1
            if ended[0]:
                return
            self.pos[(c, "prefetch")] = g
            while not c.ready(g):
                yield
            t = c.read(g)
            if t >= self.total:
                ended[0] = True
                self.pos[(c, "prefetch")] = 1 << 30
                return
            c.seen["prefetch"].append(t)
        for g in range(c.in_slots):
            yield from issue_item(g)
        it = 0
        while True:
            self.pos[(c, "epilogue")] = it
            while not c.ready(it):
                yield
            t = c.read(it)
            if t >= self.total:
                break
            c.seen["epilogue"].append(t)
            while c.tiles_mma <= c.tiles_epi:                # accumulator-full barrier
                yield
            yield
            c.tiles_epi += 1                                 # accumulator stage released
            if c.in_slots:
                yield from issue_item(it + c.in_slots)
            it += 1
        self.pos[(c, "epilogue")] = 1 << 30
        self.pos[(c, "prefetch")] = 1 << 30

    def run(self, max_steps=2_000_000):
        procs = []
        for c in self.ctas:
            if c.in_slots == 0:
                self.pos[(c, "prefetch")] = 1 << 30
            else:
                self.pos[(c, "prefetch")] = 0
            self.pos[(c, "mma")] = self.pos[(c, "epilogue")] = 0
            procs += [self.producer(c), self.mma(c), self.epilogue(c)]
        live = list(procs)
        steps = 0
        while live:
            steps += 1
            assert steps < max_steps, "no progress: deadlock in the hand-over protocol"
            p = self.rng.choice(live)
            try:
                next(p)
            except StopIteration:
                live.remove(p)
        return steps


def _check(total, n_ctas, stages, kblocks, in_slots, seed):
    sim = Sim(total, n_ctas, stages, kblocks, in_slots, random.Random(seed))
    sim.run()
    done = []
    for c in sim.ctas:
        assert c.seen["producer"] == c.seen["mma"] == c.seen["epilogue"]
        if in_slots:
            assert c.seen["prefetch"] == c.seen["epilogue"]
        done += c.seen["epilogue"]
    assert sorted(done) == list(range(total)), "every work item exactly once"


def test_handover_protocol_random_interleavings():
    rng = random.Random(0)
    for trial in range(300):
        n_ctas = rng.choice([1, 2, 3, 7])
        _check(total=rng.choice([0, 1, 2, 5, 17, 40, 131]), n_ctas=n_ctas, stages=rng.choice([2, 4, 6, 8]),
               kblocks=rng.choice([1, 1, 2, 4, 12]), in_slots=rng.choice([0, 1, 2]), seed=trial)


def test_short_tiles_with_deep_operand_ring_do_not_overrun_the_ring():
    # one k-block per item and 8 operand stages: the producer's lead over the epilogue is at its maximum (8 + 2 + published-ahead 1)
    for seed in range(50):
        _check(total=300, n_ctas=2, stages=8, kblocks=1, in_slots=2, seed=seed)
