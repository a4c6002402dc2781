"""Per-conversation state of the MCP server: search sessions, usage counters and outgoing webhooks.

Behavioural contract (`infomesh/mcp/session.py` in the reference): at most 1000 sessions that expire one
hour after their last use; counters for searches / crawls / fetches with a mean search latency; at most 20 webhook URLs,
each validated against the SSRF rules when registered and again (with DNS resolution) when fired.

Design here: sessions live in a recency-ordered map, so both capacity eviction and expiry look only at the cold end;
the latency mean is updated incrementally; counters are guarded by a thread lock because tool handlers also run in
worker threads (``asyncio.to_thread``); webhook deliveries run concurrently under a small semaphore."""
from __future__ import annotations

import asyncio
import threading
import time
from collections import Counter, OrderedDict
from dataclasses import dataclass, field

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

SESSION_CAPACITY = 1000
SESSION_IDLE_SECONDS = 3600.0
WEBHOOK_CAPACITY = 20
WEBHOOK_PARALLELISM = 8


@dataclass(slots=True)
class SearchSession:
    """What a follow-up query may refer back to."""
    last_query: str = ""
    last_results: str = ""
    updated_at: float = field(default_factory=time.time)

    def idle_for(self, now: float) -> float:
        return now - self.updated_at


class SessionStore:
    """Recency-ordered session table.  ``get_or_create`` is the only mutator: it refreshes a live session's position,
    replaces an expired one, and trims the cold end when the table is full."""

    def __init__(self, max_size: int = SESSION_CAPACITY, ttl_seconds: float = SESSION_IDLE_SECONDS):
        self._table: OrderedDict[str, SearchSession] = OrderedDict()
        self._capacity = max(1, int(max_size))
        self._idle_limit = float(ttl_seconds)

    def _expired(self, s: SearchSession, now: float) -> bool:
        return s.idle_for(now) >= self._idle_limit

    def _trim(self, now: float) -> None:
        # expired entries first (anywhere: callers may have aged a session by hand), then the least recently used
        for key in [k for k, s in self._table.items() if self._expired(s, now)]:
            del self._table[key]
        while len(self._table) >= self._capacity:
            oldest = min(self._table.items(), key=lambda kv: kv[1].updated_at)[0]
            del self._table[oldest]

    def get_or_create(self, session_id: str) -> SearchSession:
        now = time.time()
        found = self._table.get(session_id)
        if found is not None and not self._expired(found, now):
            self._table.move_to_end(session_id)
            return found
        self._table.pop(session_id, None)
        if len(self._table) >= self._capacity:
            self._trim(now)
        fresh = SearchSession(updated_at=now)
        self._table[session_id] = fresh
        return fresh

    def __len__(self) -> int:
        return len(self._table)

    def __contains__(self, session_id: object) -> bool:
        return session_id in self._table


class AnalyticsTracker:
    """Process-lifetime usage counters.  ``avg_latency_ms`` is a running mean (no sum to overflow or drift)."""

    def __init__(self):
        self._guard = threading.Lock()
        self._counts: Counter[str] = Counter()
        self.avg_latency_ms = 0.0
        self.tool_calls: dict[str, int] = {}

    # counters are exposed as read-only attributes (the status tool and the admin API read them directly)
    @property
    def total_searches(self) -> int:
        return self._counts["search"]

    @property
    def total_crawls(self) -> int:
        return self._counts["crawl"]

    @property
    def total_fetches(self) -> int:
        return self._counts["fetch"]

    async def record_search(self, latency_ms: float) -> None:
        with self._guard:
            self._counts["search"] += 1
            self.avg_latency_ms += (float(latency_ms) - self.avg_latency_ms) / self._counts["search"]

    async def record_crawl(self) -> None:
        with self._guard:
            self._counts["crawl"] += 1

    async def record_fetch(self) -> None:
        with self._guard:
            self._counts["fetch"] += 1

    def record_tool(self, name: str) -> None:
        with self._guard:
            self.tool_calls[name] = self.tool_calls.get(name, 0) + 1

    def to_dict(self) -> dict[str, object]:
        with self._guard:
            return {"total_searches": self._counts["search"], "total_crawls": self._counts["crawl"],
                    "total_fetches": self._counts["fetch"], "avg_latency_ms": round(self.avg_latency_ms, 1)}


class WebhookRegistry:
    """Outgoing notification hooks.  ``register`` answers ``None`` for success and a human-readable reason otherwise."""

    def __init__(self, max_registrations: int = WEBHOOK_CAPACITY, secret: str = ""):
        self._hooks: dict[str, None] = {}          # insertion-ordered set
        self._capacity = int(max_registrations)
        self._secret = secret

    @property
    def urls(self) -> list[str]:
        return list(self._hooks)

    def register(self, url: str) -> str | None:
        from infomesh_b200.security import SSRFError, validate_url

        try:
            validate_url(url)
        except SSRFError:
            return f"Webhook URL blocked for security: {url}"
        if url not in self._hooks:
            if len(self._hooks) >= self._capacity:
                return f"Max webhooks ({self._capacity}) reached. Unregister one first."
            self._hooks[url] = None
        return None

    def unregister(self, url: str) -> bool:
        return self._hooks.pop(url, False) is None

    async def notify(self, event: str, payload: dict[str, object]) -> int:
        """Deliver ``{event, data, timestamp}`` to every hook; returns how many answered below HTTP 400.  Each target is
        re-validated with DNS resolution right before the POST (a hostname may have been re-pointed since registration)
        and the body is HMAC-signed when a secret is configured."""
        targets = self.urls
        if not targets:
            return 0
        import httpx

        from infomesh_b200.security import validate_url
        from infomesh_b200.security_ext import sign_webhook_payload

        message: dict[str, object] = {"event": event, "data": payload, "timestamp": time.time()}
        extra = {"X-InfoMesh-Signature": sign_webhook_payload(message, self._secret)} if self._secret else {}
        gate = asyncio.Semaphore(WEBHOOK_PARALLELISM)

        async def deliver(client, target: str) -> int:
            async with gate:
                try:
                    validate_url(target, resolve_dns=True)
                    reply = await client.post(target, json=message, headers=extra)
                    return int(reply.status_code < 400)
                except Exception as exc:  # noqa: BLE001 -- SSRF refusal, DNS failure, timeout, connection error
                    logger.debug("webhook_delivery_failed", url=target, error=str(exc))
                    return 0

        async with httpx.AsyncClient(timeout=5.0) as client:
            return sum(await asyncio.gather(*(deliver(client, t) for t in targets)))
