"""CPU model of the barrier protocol of csrc/conv_kf.cu (producer / two issuers / NPART epilogue parts, mbarrier phase-parity
semantics including the false pass when a waiter is a full phase behind).  Random interleavings; reports deadlocks and reads
of an accumulator that does not hold the plane the reader expects."""
import random, sys


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self, n=1):
        self.pending -= n
        if self.pending < 0:
            raise OverflowError("too many arrivals")
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):  # try_wait.parity: true when the phase with this parity has completed
        return (self.phase & 1) != parity


def fragments(o, o_end, D):
    while o < o_end:
        col, t0 = divmod(o, D)
        left = o_end - o
        t1 = D - 1 if left >= D - t0 else t0 + left - 1
        sa, sb = max(t0 - 1, 0), min(t1 + 1, D - 1)
        o += t1 - t0 + 1
        yield col, t0, t1, sa, sb


def run(D, o, o_end, R, STAGES, NPART, seed, MW=2, NPR=1):
    """STAGES == R models the one-commit scheme of the kernel (stage = slot = g % R, the producer waits on accfull of plane g - R);
    otherwise the stage ring has its own `empty` barriers and is split between the issuers."""
    onebar = STAGES == R
    rnd = random.Random(seed)
    HS = STAGES // MW
    full = [Bar(1) for _ in range(STAGES)]
    empty = [Bar(1) for _ in range(STAGES)]
    accfull = [Bar(1) for _ in range(R)]
    accempty = [Bar(3) for _ in range(R)]
    slot_holds = [None] * R   # plane number whose MMAs have completed in the slot
    stage_holds = [None] * STAGES
    errors = []

    def producer(me=0):
        g = 0
        for col, t0, t1, sa, sb in fragments(o, o_end, D):
            for s in range(sa, sb + 1):
                if g % NPR == me:
                    m, j = g % MW, g // MW
                    if onebar:
                        st, u = g % R, g // R
                        if g >= R:
                            while not accfull[st].test((u - 1) & 1):
                                yield
                            if slot_holds[st] != g - R:
                                errors.append("producer: stage %d last multiplied %s, wanted %d" % (st, slot_holds[st], g - R))
                    else:
                        st, u = MW * (j % HS) + m, j // HS
                        while not empty[st].test((u & 1) ^ 1):
                            yield
                    stage_holds[st] = g
                    full[st].arrive()
                g += 1
                yield

    def issuer(me):
        g = 0
        for col, t0, t1, sa, sb in fragments(o, o_end, D):
            for s in range(sa, sb + 1):
                if (g % MW) == me:
                    slot, k = g % R, g // R
                    while not accempty[slot].test((k & 1) ^ 1):
                        yield
                    j = g // MW
                    st, u = (slot, k) if onebar else (MW * (j % HS) + me, j // HS)
                    while not full[st].test(u & 1):
                        yield
                    if stage_holds[st] != g:
                        errors.append("issuer %d: stage %d holds %s, wanted %d" % (me, st, stage_holds[st], g))
                    yield
                    slot_holds[slot] = g
                    if not onebar:
                        empty[st].arrive()
                    accfull[slot].arrive()
                g += 1
                yield

    def part(pt):
        g_base, oc = 0, 0
        for col, t0, t1, sa, sb in fragments(o, o_end, D):
            for t in range(t0, t1 + 1):
                if oc % NPART == pt:
                    has_m, has_p = t > 0, t < D - 1
                    g0 = g_base + t - sa
                    need = ([g0 - 1] if has_m else []) + [g0] + ([g0 + 1] if has_p else [])
                    for g in need:
                        while not accfull[g % R].test((g // R) & 1):
                            yield
                    yield
                    for g in need:
                        if slot_holds[g % R] != g:
                            errors.append("part %d out %d: slot %d holds %s, wanted %d" % (pt, oc, g % R, slot_holds[g % R], g))
                    if has_m:
                        accempty[(g0 - 1) % R].arrive(3 if t == t0 else 1)
                    accempty[g0 % R].arrive(1 + (t == t0) + (t == t1))
                    if has_p:
                        accempty[(g0 + 1) % R].arrive(3 if t == t1 else 1)
                oc += 1
                yield
            g_base += sb - sa + 1

    procs = [producer(i) for i in range(NPR)] + [issuer(i) for i in range(MW)] + [part(i) for i in range(NPART)]
    alive = list(range(len(procs)))
    idle = 0
    # adversarial schedules: every role gets a random speed (a starved issuing thread is what exposes an early parity wait)
    speed = [rnd.choice((1, 1, 1, 4, 20, 100)) for _ in procs]
    while alive:
        i = rnd.choices(alive, weights=[1.0 / speed[a] for a in alive])[0]
        state = (tuple(b.phase for b in full + empty + accfull + accempty), tuple(b.pending for b in accempty))
        try:
            next(procs[i])
        except StopIteration:
            alive.remove(i)
            idle = 0
            continue
        state2 = (tuple(b.phase for b in full + empty + accfull + accempty), tuple(b.pending for b in accempty))
        idle = idle + 1 if state == state2 else 0
        if idle > 20000:
            return "DEADLOCK (alive: %s)" % alive, errors
    return "ok", errors


if __name__ == "__main__":
    bad = 0
    perconf = {}
    # (R, NPART, MW, STAGES, NPR): the shipped kinds are conv2 (4, 2, 2, 4, 1), conv4 (4, 2, 2, 2, 1), prob (8, 4, 4, 8, 2), prob wide
    # (8, 4, 4, 8, 2) / (10, 4, 2, 10, 1), conv0 (4, 2, 2, 4, 1); (4, 4, ...) is the configuration that fails
    for R, NPART, MW, STAGES, NPR in ((4, 2, 2, 4, 1), (4, 2, 2, 2, 1), (8, 4, 4, 8, 2), (10, 4, 2, 10, 1), (4, 2, 4, 4, 2), (8, 4, 2, 8, 1),
                                     (8, 2, 2, 8, 2), (4, 1, 2, 8, 1), (6, 2, 2, 8, 1), (4, 4, 2, 8, 1)):
        for D in (1, 2, 3, 4, 5, 8):
            for trial in range(200):
                rnd = random.Random(trial)
                o = rnd.randrange(0, 3 * D)
                n = rnd.randrange(1, 40)
                try:
                    res, errs = run(D, o, o + n, R, STAGES, NPART, trial, MW, NPR)
                except OverflowError:
                    res, errs = "OVER-ARRIVAL", []
                if res != "ok" or errs:
                    bad += 1
                    perconf[(R, NPART, MW, STAGES, NPR)] = perconf.get((R, NPART, MW, STAGES, NPR), 0) + 1
                    if bad < 0:
                        print("R=%d NPART=%d D=%d range [%d,%d): %s %s" % (R, NPART, D, o, o + n, res, errs[:2]))
    print("bad cases:", bad, perconf)
