"""An executable model of clip mode's inter-CTA protocol (csrc/crt_fused_ps2.cuh) — the deadlock argument as a test.

A run of F frames of N tiles is a queue of items i = f * N + t, frame-major.  Item (f, t) may fetch its state tile once
done[t] >= f; whoever finished (f - 1, t) publishes that.  G persistent CTAs work the queue off; only R of them are resident at a
time (R = G under a cooperative launch; fewer when something else holds SMs), a waiting CTA keeps its slot.

What the kernels rely on, checked here over many random interleavings:
  * atomic counter (items handed out in order, only to running CTAs) + "no CTA blocks on a flag while it owes a publication":
    completes for every N, G, R;
  * fixed stride (item = cta + k * G): completes when every CTA is resident — and can deadlock when not (hence the cooperative
    launch, and the counter as the fallback);
  * publishing only AFTER the next item's flag wait deadlocks as soon as the grid holds two frames' worth of CTAs.
"""
import random

import pytest


def simulate(N, F, G, R, items, publish_before_wait, seed):
    """Returns True when all N * F items complete, False on deadlock.  One scheduler step = one CTA advancing one state."""
    rng = random.Random(seed)
    total = N * F
    done = [0] * N                      # done[t] = frames of tile t completed and published
    counter = [0]                       # the atomic item counter
    nxt = {c: c for c in range(G)}      # fixed stride: next item of CTA c
    waiting = list(range(G))            # CTAs not yet resident, in launch order
    running = {}                        # cta -> state
    finished = 0

    def take(c):
        if items == "counter":
            i = counter[0]; counter[0] += 1
        else:
            i = nxt[c]; nxt[c] += G
        return i if i < total else None

    def admit():
        while waiting and len(running) < R:
            c = waiting.pop(0)
            running[c] = dict(cur=take(c), owed=None, phase="top")

    admit()
    while running:
        progressed = False
        order = list(running)
        rng.shuffle(order)
        for c in order:
            s = running[c]
            if s["cur"] is None:                                   # no item left: publish what is owed and leave
                if s["owed"] is not None:
                    t, f = s["owed"]; done[t] = max(done[t], f + 1)
                del running[c]
                admit()
                progressed = True
                break
            f, t = divmod(s["cur"], N)
            if s["phase"] == "top":
                if done[t] >= f:                                   # the tile's previous frame is published: go on, publication deferred
                    s["phase"] = "work"; progressed = True; break
                if publish_before_wait and s["owed"] is not None:  # must block: publish first
                    ot, of = s["owed"]; done[ot] = max(done[ot], of + 1); s["owed"] = None
                    progressed = True; break
                continue                                           # blocked on another CTA's flag
            if s["phase"] == "work":                               # tile evaluated; the deferred publication of the previous item, then next item
                if s["owed"] is not None:
                    ot, of = s["owed"]; done[ot] = max(done[ot], of + 1)
                s["owed"] = (t, f)
                s["cur"] = take(c); s["phase"] = "top"
                finished += 1
                progressed = True
                break
        if not progressed:
            return False                                           # every resident CTA is blocked: deadlock
    return finished == total


@pytest.mark.parametrize("N,G", [(10, 6), (6, 10), (4, 8), (5, 5), (17, 4), (3, 12)])
@pytest.mark.parametrize("R", [1, 2, 1000])
def test_counter_with_publish_before_wait_never_deadlocks(N, G, R):
    for seed in range(20):
        assert simulate(N, 5, G, min(R, G), "counter", True, seed)


@pytest.mark.parametrize("N,G", [(10, 6), (6, 10), (4, 8), (5, 5), (17, 4), (3, 12)])
def test_fixed_stride_completes_when_every_cta_is_resident(N, G):
    for seed in range(20):
        assert simulate(N, 5, G, G, "stride", True, seed)


def test_fixed_stride_can_deadlock_without_full_residency():
    """10 tiles, 6 CTAs, 2 resident: CTA 0's third item (frame 1, tile 2) waits for frame 0 of tile 2 — CTA 2's first item, and
    CTA 2 never gets a slot.  The counter hands that item to a running CTA instead."""
    assert not any(simulate(10, 3, 6, 2, "stride", True, seed) for seed in range(20))
    assert all(simulate(10, 3, 6, 2, "counter", True, seed) for seed in range(20))


def test_publishing_after_the_wait_deadlocks():
    """N tiles, 2N CTAs, all resident: CTA c finishes (0, c) and takes (2, c), which waits for (1, c) — whose CTA waits for the
    publication of (0, c), which CTA c would only make after its own wait."""
    assert not any(simulate(4, 4, 8, 8, "stride", False, seed) for seed in range(20))
    assert not any(simulate(4, 4, 8, 8, "counter", False, seed) for seed in range(20))
    assert all(simulate(4, 4, 8, 8, "stride", True, seed) for seed in range(20))
