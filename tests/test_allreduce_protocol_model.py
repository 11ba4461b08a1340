"""CPU model of the deferred fused all-reduce (multibox_b200/csrc/mbx_match.cuh): outbox ring, relay bursts,
table ring, pull route, the lag.  The CUDA code is exercised on the GPU (tests/test_gpu_loopback.py,
profiles/dist_check.py); this checks the ARITHMETIC the kernels rely on -- the ring depth against lag and
burst size, the relay's skip rule -- under adversarial interleavings that a real run only produces rarely:
ranks drifting apart as far as the protocol lets them, relays that start late, stall, or never run.

Invariants (every word carries tag = step + 1, so a reader can never take a wrong value; what can go wrong is
LIVENESS -- a reader spinning for words that were overwritten before it looked):
  * a step k >= lag can always complete: every rank's words of step k - lag are in the reader's table, or
    still in the owner's outbox once the owner has completed that step;
  * a relay never overwrites table words a reader still needs;
  * the global sums every rank gets for a step are the same and are that step's sums.
Constants mirror the C side (kArRing, kRelaySteps, the defaults of mbx_allreduce_config)."""
import random

import pytest

RING = 64          # kArRing
RELAY_LIFE = 8     # kRelaySteps


class Rank:
    def __init__(self, r, world, lag):
        self.r, self.world, self.lag = r, world, lag
        self.seq = 0                                   # steps completed
        self.relayed = 0
        self.outbox = [None] * RING                    # slot -> (tag, value)
        self.table = [[None] * world for _ in range(RING)]
        self.results = {}                              # step -> global sum this rank computed for it
        self.pulls = 0


def value(r, step):
    return (r + 1) * 1000003 + step * 7


def words_in_table(me, step):
    row = me.table[step % RING]
    return all(w is not None and w[0] == step + 1 for w in row)


def can_complete(me, ranks, lag):
    """The last CTA of `me`'s next step: table, else pull -- the pull spins until every owner has completed the step."""
    k = me.seq
    if k < lag:
        return True
    s = k - lag
    if words_in_table(me, s):
        return True
    return all(o.seq > s for o in ranks)


def complete(me, ranks, lag):
    k = me.seq
    me.outbox[k % RING] = (k + 1, value(me.r, k))      # finalize_losses: own outbox first
    if k >= lag:
        s = k - lag
        if words_in_table(me, s):
            tot = sum(w[1] for w in me.table[s % RING])
        else:
            me.pulls += 1
            tot = 0
            for o in ranks:                            # ar_pull_warp: the owner's outbox must STILL hold step s
                w = o.outbox[s % RING]
                assert w is not None and w[0] == s + 1, \
                    "outbox slot of step %d overwritten (owner at %d, reader at %d)" % (s, o.seq, k)
                tot += w[1]
        me.results[s] = tot
    me.seq = k + 1


def relay_burst(me, ranks, batch, idle):
    """One iteration of mbx_allreduce_relay_kernel; returns the number of steps forwarded."""
    s = me.relayed
    if me.seq - s >= RING // 2:                        # stale: consumed everywhere
        s = me.seq - RING // 2 + 1
    hi = s + batch - 1
    w = me.outbox[hi % RING]
    ready = w is not None and w[0] == hi + 1
    if not (ready or idle):
        me.relayed = max(me.relayed, s)
        return 0
    sent = 0
    for q in range(s, hi + 1):
        w = me.outbox[q % RING]
        if w is None or w[0] != q + 1:
            break
        for o in ranks:
            old = o.table[q % RING][me.r]
            # the slot may hold an older step's words; never a NEWER one (that would be a stale relay)
            assert old is None or old[0] <= q + 1, "relay of rank %d overwrote newer words" % me.r
            # ... and the reader no longer needs the words it replaces (it reads step t at its step t + lag)
            if old is not None and old[0] != q + 1:
                assert o.seq > (old[0] - 1) + o.lag, "relay of rank %d overwrote words rank %d still needs" % (me.r, o.r)
            o.table[q % RING][me.r] = w
        sent += 1
    me.relayed = s + sent
    return sent


@pytest.mark.parametrize("world,lag,batch", [(2, 12, 8), (8, 12, 8), (5, 1, 1), (8, 4, 1), (3, 8, 4), (8, 24, 8)])
@pytest.mark.parametrize("relay_mode", ["prompt", "late", "absent", "random"])
def test_ring_depth_and_liveness(world, lag, batch, relay_mode):
    assert 2 * lag + batch < RING      # mbx_allreduce_config's rule
    rng = random.Random(1234 + world * 100 + lag * 10 + batch)
    ranks = [Rank(r, world, lag) for r in range(world)]
    target = 400
    relay_alive = [relay_mode != "absent"] * world
    relay_left = [RELAY_LIFE] * world
    stuck = 0
    while min(r.seq for r in ranks) < target:
        progressed = False
        # an adversarial scheduler: prefer the rank that is AHEAD (maximises drift) most of the time
        order = sorted(ranks, key=lambda x: -x.seq) if rng.random() < 0.7 else rng.sample(ranks, world)
        for me in order:
            if me.seq < target + 2 * lag and can_complete(me, ranks, lag):
                complete(me, ranks, lag)
                progressed = True
                break
        for me in ranks:                               # relays
            if relay_mode == "absent":
                continue
            if relay_mode == "late" and rng.random() < 0.9:
                continue
            if relay_mode == "random":
                if rng.random() < 0.02:
                    relay_alive[me.r] = not relay_alive[me.r]
                if not relay_alive[me.r]:
                    continue
            idle = rng.random() < 0.05
            n = relay_burst(me, ranks, batch, idle)
            relay_left[me.r] -= n
            if relay_left[me.r] <= 0 or idle:          # a relay's life; the host side launched the next one
                relay_left[me.r] = RELAY_LIFE
            progressed = progressed or n > 0
        stuck = 0 if progressed else stuck + 1
        assert stuck < 50, "deadlock: seq=%r" % [r.seq for r in ranks]
        # ranks are never more than `lag` steps apart
        assert max(r.seq for r in ranks) - min(r.seq for r in ranks) <= lag + 1
    # every rank computed the same, correct global sum for every step it completed a reduction for
    for s in range(target - lag):
        want = sum(value(r, s) for r in range(world))
        for me in ranks:
            assert me.results[s] == want, (me.r, s)
    if relay_mode == "absent":
        assert all(me.pulls >= target - lag for me in ranks)      # everything went the pull route
    if relay_mode == "prompt" and batch == 1:
        assert all(me.pulls <= lag + 2 for me in ranks)           # the relay keeps up: (almost) nothing is pulled


def test_model_trips_when_the_ring_is_too_small(monkeypatch):
    """The model has teeth: with a 16-deep ring and lag 9 (2 * lag > ring) words are overwritten before their
    readers look, in every relay mode."""
    import sys
    mod = sys.modules[__name__]
    monkeypatch.setattr(mod, "RING", 16)
    body = test_ring_depth_and_liveness.__wrapped__ if hasattr(test_ring_depth_and_liveness, "__wrapped__") \
        else test_ring_depth_and_liveness
    for mode in ("prompt", "late", "absent", "random"):
        with pytest.raises(AssertionError, match="overwr|rule"):
            _run_without_rule(body, 8, 9, 1, mode)


def _run_without_rule(body, world, lag, batch, mode):
    # the first statement of the test is mbx_allreduce_config's own admission rule; the point here is what
    # happens when it is ignored
    import inspect
    import textwrap
    src = textwrap.dedent(inspect.getsource(body))
    src = src[src.index("def "):].replace("    assert 2 * lag + batch < RING      # mbx_allreduce_config's rule\n", "")
    ns = dict(globals())
    exec(compile(src, "<model>", "exec"), ns)      # noqa: S102  (test-only: re-run the model without its guard)
    ns["test_ring_depth_and_liveness"](world, lag, batch, mode)
