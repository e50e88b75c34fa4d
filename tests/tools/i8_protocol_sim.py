"""Discrete-event model of the barrier protocol of predict_i8_kernel (csrc/predict_i8.cu): one producer, two MMA
issuers taking the stages in turn, the epilogue, one in-order tensor pipe, mbarriers with ONE parity bit.

Test infrastructure (imported by tests/test_host_logic.py only).  It exists because two protocol errors of the
bring-up were found (or explained) with it and not on the device:
  * a segment with a single stage lets the issuer that has no stage in it run a whole segment ahead, where a parity
    wait can no longer tell the phases apart  ->  every segment has two stages (an empty second one if need be);
  * copies of different stages complete out of order; with an odd ring an issuer meets a slot at every other use
    only, and the parity test for "my use" also passes when the use in between has not landed yet  ->  the full
    barriers are indexed n % (issuers x stages) so that each belongs to one issuer (nfull argument);
  * the producer emits the stages of a chunk lane-parallel, each lane as soon as the ring slot of its stage is free; a
    lane many stages ahead would read "free" off the slot's barrier for the same reason (one parity bit)  ->  a lane may
    look at its barrier only when the stage one ring earlier has been emitted: a window of `stages` stages past the
    emitted prefix (window argument; None = any lane may look, which the model shows to be wrong).
run(seed, segments, stages, nfull) returns "ok" or a description of the violation / deadlock."""
import random


class MBar:
    def __init__(self, count, name):
        self.count, self.pending, self.bit, self.name = count, count, 0, name

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, (self.name, "over-arrival")
        if self.pending == 0:
            self.bit ^= 1
            self.pending = self.count

    def test(self, parity):            # mbarrier.test_wait.parity: has the phase of this parity completed?
        return self.bit != parity


def stage_list(segments):
    """(segment, flags) per stage as the producer emits them: at least two stages per segment."""
    out = []
    for c, nl in enumerate(segments):
        nl = max(nl, 2)
        for i in range(nl):
            f = set()
            if i == 0:
                f.add("FIRST")
            if i == 1:
                f.add("SECOND")
            if i == nl - 2:
                f.add("PENULT")
            if i == nl - 1:
                f.add("LAST")
            out.append((c, f))
    out[-1][1].add("FINAL")
    out.append((None, {"EXIT"}))
    return out


def run(seed, segments, stages=5, nfull=None, weights=None, min_two=True, chunk=8, window="stages"):
    rnd = random.Random(seed)
    nfull = nfull or 2 * stages
    full = [MBar(1, "full%d" % i) for i in range(nfull)]
    empty = [MBar(1, "empty%d" % i) for i in range(stages)]
    acc_full, acc_empty, first_done = MBar(2, "acc_full"), MBar(1, "acc_empty"), MBar(1, "first_done")
    hdr = [None] * stages
    sl = stage_list(segments if min_two else segments)
    pipe, landing = [], []                      # tensor pipe (in order); copies in flight (any order)
    acc = {"seg": -1, "drained": True}

    in_ring = {}                                 # slot -> stage whose data the slot holds, until its products have completed

    def producer():
        # chunks of `chunk` stages are emitted lane-parallel: in every round each lane that is inside the window and
        # finds its slot's barrier in the "previous use is over" state emits its stage
        win = stages if window == "stages" else (10 ** 9 if window is None else window)
        for base in range(0, len(sl), chunk):
            todo = list(range(base, min(base + chunk, len(sl))))
            while todo:
                prefix = todo[0]                 # every stage below has been emitted
                for n in list(todo):
                    if n >= prefix + win or rnd.random() < 0.3:
                        continue
                    rs = n % stages
                    if n < stages or empty[rs].test(((n // stages) & 1) ^ 1):
                        assert rs not in in_ring, ("stage %d written over stage %d, which is still in the ring" % (n, in_ring[rs]))
                        in_ring[rs] = n
                        hdr[rs] = sl[n] + (n,)
                        landing.append(n % nfull)    # the copy lands (completes the full barrier) some time later
                        todo.remove(n)
                yield "producer: ring slots of stages %d.." % todo[0] if todo else None

    def copies():
        while True:
            if landing and rnd.random() < 0.5:
                full[landing.pop(rnd.randrange(len(landing)))].arrive()
            yield None

    def issuer(w):
        n = w
        while True:
            while not full[n % nfull].test((n // nfull) & 1):
                yield "issuer %d: full of stage %d" % (w, n)
            c, f, nn = hdr[n % stages]
            assert nn == n, ("issuer %d took the header of stage %d for stage %d" % (w, nn, n))
            if "EXIT" in f:
                return
            if "FIRST" in f and c > 0:
                while not acc_empty.test((c - 1) & 1):
                    yield "issuer %d: acc_empty at stage %d" % (w, n)
            if "SECOND" in f:
                while not first_done.test(c & 1):
                    yield "issuer %d: first_done at stage %d" % (w, n)
            for q in range(3 if "FIRST" in f else rnd.randint(0, 3)):
                pipe.append(["mma", w, c, "FIRST" in f and q == 0])
                if rnd.random() < 0.5:
                    yield None                  # the other issuer's products may slip in between
            pipe.append(["commit", w, ("empty", n % stages)])
            if "FIRST" in f:
                first_done.arrive()
            if "LAST" in f or "PENULT" in f:
                pipe.append(["commit", w, ("acc_full", c)])
            yield None
            if "FINAL" in f:
                return
            n += 2

    def tensor():
        while True:
            fired = True
            while fired:                        # a commit fires once no MMA of its issuer precedes it
                fired, seen = False, set()
                for i, it in enumerate(pipe):
                    if it[0] == "mma":
                        seen.add(it[1])
                    elif it[1] not in seen:
                        pipe.pop(i)
                        if it[2][0] == "empty":
                            in_ring.pop(it[2][1], None)
                        (empty[it[2][1]] if it[2][0] == "empty" else acc_full).arrive()
                        fired = True
                        break
            if pipe and pipe[0][0] == "mma" and rnd.random() < 0.7:
                _, _, c, first = pipe.pop(0)
                if first:
                    assert acc["seg"] == c - 1 and acc["drained"], ("accumulators overwritten before the drain", c)
                    acc["seg"], acc["drained"] = c, False
                else:
                    assert acc["seg"] == c and not acc["drained"], ("product added to the wrong segment", c, acc["seg"])
            yield None

    def epilogue():
        ph = 0
        for c in range(len(segments)):
            while not acc_full.test(ph):
                yield "epilogue: acc_full of segment %d" % c
            ph ^= 1
            assert acc["seg"] == c and not any(it[0] == "mma" and it[2] == c for it in pipe), ("drained early", c)
            acc["drained"] = True
            acc_empty.arrive()
            yield None

    procs = {"producer": producer(), "copies": copies(), "issuer0": issuer(0), "issuer1": issuer(1),
             "epilogue": epilogue(), "tensor": tensor()}
    weights = weights or {}
    waiting, stuck = {}, 0
    workers = ("producer", "issuer0", "issuer1", "epilogue")
    try:
        while any(k in procs for k in workers):
            names = list(procs)
            k = rnd.choices(names, [weights.get(x, 1.0) for x in names])[0]
            try:
                waiting[k] = next(procs[k])
            except StopIteration:
                del procs[k]
                waiting.pop(k, None)
                continue
            if not pipe and not landing and all(waiting.get(x) for x in workers if x in procs):
                stuck += 1
                if stuck > 20000:
                    return "deadlock: " + "; ".join(str(waiting[x]) for x in workers if x in procs)
            else:
                stuck = 0
    except AssertionError as e:
        return "violation: %s" % (e,)
    return "ok"


if __name__ == "__main__":
    import sys
    bad = 0
    for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2000):
        r = random.Random(seed)
        segs = [r.choice([1, 1, 2, 2, 3, 4, 7, 12]) for _ in range(r.randint(1, 10))]
        w = {k: r.choice([0.05, 0.3, 1, 4]) for k in ("producer", "copies", "issuer0", "issuer1", "epilogue", "tensor")}
        for stages in (5, 4):
            res = run(seed, segs, stages=stages, weights=w)
            if res != "ok":
                print(seed, segs, stages, res)
                bad += 1
    print("violations:", bad)
