"""Model check of the weight-gradient exchange protocol of csrc/peer.cu (no GPU, no native code: a small state
machine explored exhaustively).

What the kernels do per exchange (epoch e) on every rank r, in stream order:
    push    W(e): write this rank's slices into its OWN slot e & 1
    finish  P(e): publish -- store e + 1 into flag[q][r] on every rank q (one store per peer)
            A(e): wait until flag[r][src] >= e + 1 for every source
            R(e): read slot e & 1 of every rank (its own locally, the others through the peer mapping)
Claims checked over ALL interleavings of the ranks' steps: (1) every read returns the data of the epoch it is meant
for -- in particular a slot is never overwritten (by the push of e + 2) while a peer may still read it; (2) no deadlock.
Negative controls show the checker can fail: a single slot instead of two, and publishing before the slices are written."""
import sys

import pytest


def _program(rank, world, epochs, slots, publish_first):
    prog = []
    for e in range(epochs):
        w = [("W", e, e % slots)]
        p = [("P", e, q) for q in range(world)]
        prog += (p + w) if publish_first else (w + p)
        prog.append(("A", e, None))
        prog += [("R", e, src) for src in range(world)]
    return prog


def explore(world, epochs, slots=2, publish_first=False):
    """DFS over all interleavings; returns (violation or None, states visited)"""
    progs = [_program(r, world, epochs, slots, publish_first) for r in range(world)]
    start = (tuple(0 for _ in range(world)),                                  # program counters
             tuple(tuple(None for _ in range(slots)) for _ in range(world)),  # slot[r][s] = epoch of the data in it
             tuple(tuple(0 for _ in range(world)) for _ in range(world)))     # flag[q][src]
    seen, stack = {start}, [start]
    while stack:
        pcs, slot, flag = stack.pop()
        moved = False
        for r in range(world):
            if pcs[r] >= len(progs[r]):
                continue
            op, e, arg = progs[r][pcs[r]]
            nslot, nflag = slot, flag
            if op == "W":
                row = list(slot[r]); row[arg] = e
                nslot = slot[:r] + (tuple(row),) + slot[r + 1:]
            elif op == "P":
                row = list(flag[arg]); row[r] = e + 1
                nflag = flag[:arg] + (tuple(row),) + flag[arg + 1:]
            elif op == "A":
                if any(flag[r][src] < e + 1 for src in range(world)):
                    continue                                                  # blocked
            elif op == "R":
                got = slot[arg][e % slots]
                if got != e:
                    return f"rank {r} reads rank {arg}'s slot for epoch {e} and finds epoch {got}", len(seen)
            moved = True
            nxt = (pcs[:r] + (pcs[r] + 1,) + pcs[r + 1:], nslot, nflag)
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
        if not moved and any(pcs[r] < len(progs[r]) for r in range(world)):
            return f"deadlock at program counters {pcs}", len(seen)
    return None, len(seen)


@pytest.mark.parametrize("world,epochs", [(2, 5), (3, 4), (4, 3)])
def test_two_slots_are_enough_and_nobody_deadlocks(world, epochs):
    bad, states = explore(world, epochs)
    assert bad is None, bad
    assert states > 100                                                       # the search really interleaved


def test_the_checker_finds_the_overwrite_with_a_single_slot():
    bad, _ = explore(2, 3, slots=1)
    assert bad is not None and "finds epoch" in bad


def test_the_checker_finds_a_publish_that_precedes_the_data():
    bad, _ = explore(2, 2, publish_first=True)
    assert bad is not None and "finds epoch" in bad


if __name__ == "__main__":
    for w, e in ((2, 5), (3, 4), (4, 3)):
        print(w, e, explore(w, e))
    sys.exit(0)
