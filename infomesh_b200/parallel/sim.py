"""Single-process model of the peer-memory flag protocol (csrc/comm/symm.cu, parallel/symm.py) for CPU testing.

The GPU kernels signal through cumulative arrival counters that live in every rank's symmetric heap and a private
``{use, done}`` pair advanced by the consuming kernel's last CTA; payload slots are double-buffered on ``use & 1``.
The correctness argument ("a peer can never overwrite a slot that has not been read yet, because finishing use s+1
needs this rank's use s+1 arrivals, which are stream-ordered after this rank's use s reads") is easy to get subtly
wrong, and 8 GPUs are not always at hand.  This module executes the same protocol with plain Python lists, one
generator per rank, and lets a scheduler interleave the ranks' micro-steps arbitrarily (every ``yield`` is a point where
another rank may run), so tests can throw thousands of random or adversarial schedules at it -- the analogue of the
reference's "two libp2p hosts in one process" trick.

Modelled faithfully: per-source counters only ever grow; a use with ``ctas`` producer CTAs is complete when
``flag >= (use + 1) * ctas``; each CTA pushes its slice to every peer and then bumps the counter (release); the consumer
acquires the counters, reads, and only then does the last CTA out advance ``use``.  ``buffers=1`` switches the double
buffering off so that tests can show the checker actually catches the overwrite it is meant to rule out."""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import Callable, Iterator


@dataclass
class RankHeap:
    world: int
    ctas: int
    buffers: int = 2
    slots: list = field(init=False)        # [buffers][src][cta] -> payload written by CTA `cta` of rank `src`
    flags: list = field(init=False)        # [src] cumulative arrivals from rank `src`
    use: int = 0                           # private channel state, advanced by the consumer

    def __post_init__(self) -> None:
        self.slots = [[[None] * self.ctas for _ in range(self.world)] for _ in range(self.buffers)]
        self.flags = [0] * self.world


class Violation(AssertionError):
    pass


def allgather_rank(rank: int, heaps: list[RankHeap], n_uses: int, log: list) -> Iterator[None]:
    """One rank running ``n_uses`` back-to-back push all-gathers; payload of use u from rank r, CTA c is ``(u, r, c)``."""
    me = heaps[rank]
    W, C = me.world, me.ctas
    for _ in range(n_uses):
        use = me.use                                            # every CTA reads the state word at kernel start
        par = use % me.buffers
        for c in range(C):                                      # CTAs run in some order; model them sequentially
            for pp in range(W):
                peer = heaps[(rank + pp) % W]                   # staggered destinations, own rank included
                peer.slots[par][rank][c] = (use, rank, c)       # NVLink store of this CTA's slice
                yield
            for pp in range(W):
                heaps[(rank + pp) % W].flags[rank] += 1         # red.release.sys after the CTA's stores
                yield
        target = (use + 1) * C
        while any(me.flags[src] < target for src in range(W)):  # ld.acquire.sys spin
            yield
        for src in range(W):                                    # copy-out of the gathered slots
            for c in range(C):
                got = me.slots[par][src][c]
                if got != (use, src, c):
                    raise Violation(f"rank {rank} use {use}: slot[{src}][{c}] holds {got}")
                yield
        log.append((rank, use))
        me.use = use + 1                                        # last CTA out advances the channel


def run(world: int, n_uses: int, *, ctas: int = 2, buffers: int = 2, seed: int = 0,
        pick: Callable[[random.Random, list[int]], int] | None = None, max_steps: int = 2_000_000) -> list:
    """Interleave the ranks until all finish.  ``pick(rng, runnable)`` chooses who moves next (default: uniform)."""
    rng = random.Random(seed)
    heaps = [RankHeap(world, ctas, buffers) for _ in range(world)]
    log: list = []
    gens = {r: allgather_rank(r, heaps, n_uses, log) for r in range(world)}
    steps = 0
    while gens:
        runnable = sorted(gens)
        r = pick(rng, runnable) if pick else rng.choice(runnable)
        try:
            next(gens[r])
        except StopIteration:
            del gens[r]
        steps += 1
        if steps > max_steps:
            raise RuntimeError("schedule did not terminate (livelock in the model?)")
    return log


def favour(rank: int, weight: int = 20) -> Callable[[random.Random, list[int]], int]:
    """Adversarial scheduler: let one rank race ahead of the others as far as the protocol allows."""
    def pick(rng: random.Random, runnable: list[int]) -> int:
        if rank in runnable and rng.randrange(weight):
            return rank
        return rng.choice(runnable)
    return pick
