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


# ---------------------------------------------------------------------------------------------------------------------
# Cluster row-statistics exchange of the fused GEMM + LayerNorm epilogue (csrc/gemm/gemm_mxf8.cu, `LNF`)
# ---------------------------------------------------------------------------------------------------------------------
# Per tile every epilogue warp of every CTA of the cluster sends its partial row statistics to ALL CTAs with st.async: the
# data lands in the receiver's slot and completes transaction bytes on the receiver's mbarrier in one step.  The receiver's
# warp 0 posts `arrive.expect_tx(n_cta * warps)` once per tile (the barrier's arrival count is 1); slots and barriers are
# double-buffered on `tile & 1`, a warp waits for phase `tile >> 1` of its buffer's barrier and then reads every slot.
# What has to hold under ANY interleaving of warps and CTAs: a phase never completes before all of its data has landed (even
# when data overtakes the expect_tx and the transaction count goes negative), and no slot is overwritten before every local
# warp has read it.

@dataclass
class TxBarrier:
    """mbarrier with an arrival count of 1 and a transaction count."""
    phase: int = 0
    pending: int = 1
    tx: int = 0

    def _maybe_complete(self) -> None:
        if self.pending == 0 and self.tx == 0:
            self.phase, self.pending = self.phase + 1, 1

    def expect_tx(self, units: int) -> None:        # mbarrier.arrive.expect_tx: one arrival + `units` expected
        self.tx += units
        self.pending -= 1
        if self.pending < 0:
            raise Violation("two expect_tx arrivals in one phase")
        self._maybe_complete()

    def complete_tx(self, units: int = 1) -> None:  # the st.async data landing
        self.tx -= units
        self._maybe_complete()


@dataclass
class ClusterCta:
    n_cta: int
    warps: int
    buffers: int = 2
    slots: list = field(init=False)                 # [buffer][src cta][src warp]
    bars: list = field(init=False)

    def __post_init__(self) -> None:
        self.slots = [[[None] * self.warps for _ in range(self.n_cta)] for _ in range(self.buffers)]
        self.bars = [TxBarrier() for _ in range(self.buffers)]


def stats_exchange_warp(cta: int, warp: int, ctas: list[ClusterCta], n_tiles: int, log: list) -> Iterator[None]:
    me = ctas[cta]
    for tile in range(n_tiles):
        buf = tile % me.buffers
        if warp == 0:
            me.bars[buf].expect_tx(me.n_cta * me.warps)
            yield
        for p in range(me.n_cta):                    # st.async to every CTA of the cluster, own CTA included
            ctas[p].slots[buf][cta][warp] = (tile, cta, warp)
            ctas[p].bars[buf].complete_tx(1)
            yield
        want_phase = tile // me.buffers + 1          # the (tile // buffers)-th completion of this buffer's barrier
        while me.bars[buf].phase < want_phase:       # mbarrier.try_wait.parity spin
            yield
        for src in range(me.n_cta):
            for w in range(me.warps):
                got = me.slots[buf][src][w]
                if got != (tile, src, w):
                    raise Violation(f"cta {cta} warp {warp} tile {tile}: slot[{src}][{w}] holds {got}")
                yield
        log.append((cta, warp, tile))


def run_cluster(n_cta: int, warps: int, n_tiles: int, *, buffers: int = 2, seed: int = 0,
                pick: Callable[[random.Random, list], object] | None = None, max_steps: int = 4_000_000) -> list:
    """Interleave every (CTA, warp) of one cluster over ``n_tiles`` tiles; returns the completion log."""
    rng = random.Random(seed)
    ctas = [ClusterCta(n_cta, warps, buffers) for _ in range(n_cta)]
    log: list = []
    gens = {(c, w): stats_exchange_warp(c, w, ctas, n_tiles, log) for c in range(n_cta) for w in range(warps)}
    steps = 0
    while gens:
        runnable = sorted(gens)
        who = pick(rng, runnable) if pick else rng.choice(runnable)
        try:
            next(gens[who])
        except StopIteration:
            del gens[who]
        steps += 1
        if steps > max_steps:
            raise RuntimeError("cluster schedule did not terminate (deadlock in the model?)")
    return log


def favour_cta(cta: int, weight: int = 30) -> Callable[[random.Random, list], object]:
    """Let every warp of one CTA race ahead of the rest of the cluster."""
    def pick(rng: random.Random, runnable: list):
        mine = [x for x in runnable if x[0] == cta]
        return rng.choice(mine) if mine and rng.randrange(weight) else rng.choice(runnable)
    return pick


def starve_warp0(weight: int = 30) -> Callable[[random.Random, list], object]:
    """Delay the warps that post expect_tx, so that data overtakes it and the transaction count goes negative."""
    def pick(rng: random.Random, runnable: list):
        others = [x for x in runnable if x[1] != 0]
        return rng.choice(others) if others and rng.randrange(weight) else rng.choice(runnable)
    return pick
