"""Query routing: keywords -> DHT pointers -> peer scores -> latency-aware top-5 fan-out of SEARCH_REQUESTs with an
adaptive per-peer timeout -> merged, score-sorted results; plus the responder side
(reference infomesh/p2p/routing.py:42-435).

Differences by design: asyncio ``gather`` replaces the trio nursery + memory channel; keyword lookups run
concurrently; the responder enforces :class:`NodeLoadGuard` (declared but unwired in the reference);
duplicate URLs returned by several peers are collapsed to the best-scoring copy.
"""
from __future__ import annotations

import asyncio
import time
from collections import deque
from dataclasses import dataclass, field
from typing import Any, Awaitable, Callable

from infomesh_b200.p2p.load_guard import NodeLoadGuard
from infomesh_b200.p2p.peer_profile import PeerProfileTracker
from infomesh_b200.p2p.protocol import MessageType, SearchRequest, SearchResponse, dataclass_to_payload
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

SEARCH_TIMEOUT_MS = 5000
MAX_FANOUT = 5
MAX_RESULTS_PER_PEER = 20
HEDGE_TIMEOUT_FRACTION = 0.5


@dataclass
class RoutingStats:
    queries_routed: int = 0
    queries_local_only: int = 0
    peers_contacted: int = 0
    peers_responded: int = 0
    peers_timed_out: int = 0
    avg_response_ms: float = 0.0
    _response_times: deque = field(default_factory=lambda: deque(maxlen=10_000), repr=False)

    def record_response(self, elapsed_ms: float) -> None:
        self.peers_responded += 1
        self._response_times.append(elapsed_ms)
        self.avg_response_ms = sum(self._response_times) / len(self._response_times)


@dataclass(frozen=True)
class RemoteSearchResult:
    url: str
    title: str
    snippet: str
    score: float
    peer_id: str
    doc_id: int
    elapsed_ms: float = 0.0


def _payload_str(v: object, *, default: str = "") -> str:
    return v if isinstance(v, str) else default


def _number(v: object, kinds: tuple[type, ...], parse: Callable[[Any], Any]) -> Any | None:
    """``parse(v)`` for a value of one of ``kinds`` (booleans never count as numbers); None when it is not, or does not parse."""
    if isinstance(v, bool) or not isinstance(v, kinds):
        return None
    try:
        return parse(v)
    except (TypeError, ValueError, OverflowError):
        return None


def _payload_int(v: object, *, default: int = 0) -> int:
    """Integers and integer strings (peers on other stacks serialise numbers as text); floats are not silently truncated."""
    n = _number(v, (int, str), int)
    return default if n is None else n


def _payload_float(v: object, *, default: float = 0.0) -> float:
    """Finite numbers and numeric strings; NaN and the infinities fall back to the default."""
    f = _number(v, (int, float, str), float)
    return default if f is None or f != f or f in (float("inf"), -float("inf")) else f


SendFn = Callable[[str, MessageType, dict[str, Any], float], Awaitable[tuple[MessageType, dict[str, Any]] | None]]


class QueryRouter:
    """``send`` is ``async (peer_id, type, payload, timeout_s) -> (type, payload) | None`` — the node binds it to its
    transport + address book; tests bind it to in-memory peers."""

    def __init__(self, send: SendFn, dht: Any, local_peer_id: str, *, connected_peers: Callable[[], list[str]] | None = None,
                 max_fanout: int = MAX_FANOUT, timeout_ms: int = SEARCH_TIMEOUT_MS,
                 profile_tracker: PeerProfileTracker | None = None, load_guard: NodeLoadGuard | None = None):
        self._send, self._dht, self._peer_id = send, dht, local_peer_id
        self._connected = connected_peers or (lambda: [])
        self._max_fanout, self._timeout_ms = max_fanout, timeout_ms
        self._profiles = profile_tracker or PeerProfileTracker()
        self._guard = load_guard or NodeLoadGuard()
        self._stats = RoutingStats()
        self._seq = 0

    @property
    def stats(self) -> RoutingStats:
        return self._stats

    @property
    def profile_tracker(self) -> PeerProfileTracker:
        return self._profiles

    @property
    def load_guard(self) -> NodeLoadGuard:
        return self._guard

    # ------------------------------------------------------------------ requester
    async def _candidate_scores(self, keywords: list[str]) -> dict[str, float]:
        scores: dict[str, float] = {}
        lists = await asyncio.gather(*(self._dht.query_keyword(kw) for kw in keywords), return_exceptions=True)
        for ptrs in lists:
            if not isinstance(ptrs, list):
                continue
            for p in ptrs:
                pid = _payload_str(p.get("peer_id")) if isinstance(p, dict) else ""
                if pid and pid != self._peer_id:
                    scores[pid] = scores.get(pid, 0.0) + _payload_float(p.get("score"), default=0.5)
        return scores

    async def route_query(self, query: str, keywords: list[str], limit: int = 10) -> list[RemoteSearchResult]:
        self._stats.queries_routed += 1
        if limit <= 0:
            return []
        scores = await self._candidate_scores(keywords)
        if not scores:
            scores = {pid: 0.1 for pid in self._connected() if pid and pid != self._peer_id}
            if not scores:
                self._stats.queries_local_only += 1
                return []
        ranked = sorted(scores, key=scores.get, reverse=True)[: self._max_fanout * 2]
        targets = self._profiles.rank_by_latency(ranked, diversity=True)[: self._max_fanout]
        self._stats.peers_contacted += len(targets)
        self._seq += 1
        req = dataclass_to_payload(SearchRequest(query=query, keywords=list(keywords), limit=min(limit, MAX_RESULTS_PER_PEER),
                                                 request_id=f"{self._peer_id}:{time.time():.0f}:{self._seq}"))

        async def ask(pid: str) -> list[RemoteSearchResult]:
            budget_ms = self._profiles.adaptive_timeout(pid, base_ms=float(self._timeout_ms))
            t0 = time.monotonic()
            try:
                reply = await asyncio.wait_for(self._send(pid, MessageType.SEARCH_REQUEST, req, budget_ms / 1000),
                                               timeout=budget_ms / 1000)
            except Exception as exc:  # noqa: BLE001 — timeout, refused connection, bad frame: all "peer failed"
                self._stats.peers_timed_out += 1
                self._profiles.record(pid, (time.monotonic() - t0) * 1000, success=False)
                logger.debug("peer_query_failed", peer_id=pid[:16], error=type(exc).__name__)
                return []
            ms = (time.monotonic() - t0) * 1000
            if not reply or reply[0] != MessageType.SEARCH_RESPONSE:
                self._profiles.record(pid, ms, success=False)
                return []
            self._stats.record_response(ms)
            self._profiles.record(pid, ms, success=True)
            out = []
            rows = reply[1].get("results", [])
            for r in rows[:MAX_RESULTS_PER_PEER] if isinstance(rows, list) else []:
                if isinstance(r, dict) and _payload_str(r.get("url")):
                    out.append(RemoteSearchResult(_payload_str(r.get("url")), _payload_str(r.get("title")),
                                                  _payload_str(r.get("snippet")), _payload_float(r.get("score")), pid,
                                                  _payload_int(r.get("doc_id")), ms))
            return out

        best: dict[str, RemoteSearchResult] = {}
        for batch in await asyncio.gather(*(ask(p) for p in targets)):
            for r in batch:
                if r.url not in best or r.score > best[r.url].score:
                    best[r.url] = r
        return sorted(best.values(), key=lambda r: r.score, reverse=True)[:limit]

    # ------------------------------------------------------------------ responder
    async def handle_search_request(self, payload: dict[str, Any], local_search_fn: Callable[[str, int], Awaitable[list[dict]]],
                                    *, requester: str = "") -> tuple[MessageType, dict[str, Any]]:
        if not self._guard.try_acquire(requester):
            return MessageType.ERROR, dict(self._guard.get_reject_info())
        try:
            query = _payload_str(payload.get("query"))
            limit = min(max(_payload_int(payload.get("limit"), default=10), 1), 100)
            t0 = time.monotonic()
            results = await local_search_fn(query, limit) if query.strip() else []
            resp = SearchResponse(request_id=_payload_str(payload.get("request_id")), results=list(results)[:limit],
                                  peer_id=self._peer_id, elapsed_ms=(time.monotonic() - t0) * 1000)
            return MessageType.SEARCH_RESPONSE, dataclass_to_payload(resp)
        finally:
            self._guard.release(requester)
