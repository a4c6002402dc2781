"""Conversation sessions (bounded, 1 h TTL), in-memory analytics counters and the SSRF-checked webhook registry used
by the MCP server (reference infomesh/mcp/session.py:20-241)."""
from __future__ import annotations

import asyncio
import time

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_SESSION_MAX_SIZE = 1000
_SESSION_TTL_SECONDS = 3600.0
_WEBHOOK_MAX_REGISTRATIONS = 20


class SearchSession:
    __slots__ = ("last_query", "last_results", "updated_at")

    def __init__(self):
        self.last_query, self.last_results, self.updated_at = "", "", 0.0


class SessionStore:
    __slots__ = ("_sessions", "_max_size", "_ttl")

    def __init__(self, max_size: int = _SESSION_MAX_SIZE, ttl_seconds: float = _SESSION_TTL_SECONDS):
        self._sessions: dict[str, SearchSession] = {}
        self._max_size, self._ttl = max_size, ttl_seconds

    def get_or_create(self, session_id: str) -> SearchSession:
        now = time.time()
        cur = self._sessions.get(session_id)
        if cur is not None:
            if now - cur.updated_at < self._ttl:
                return cur
            del self._sessions[session_id]
        if len(self._sessions) >= self._max_size:
            self._evict(now)
        while len(self._sessions) >= self._max_size:
            del self._sessions[min(self._sessions, key=lambda k: self._sessions[k].updated_at)]
        s = self._sessions[session_id] = SearchSession()
        s.updated_at = now
        return s

    def _evict(self, now: float) -> None:
        for k in [k for k, v in self._sessions.items() if now - v.updated_at >= self._ttl]:
            del self._sessions[k]

    def __len__(self) -> int:
        return len(self._sessions)

    def __contains__(self, session_id: str) -> bool:
        return session_id in self._sessions


class AnalyticsTracker:
    __slots__ = ("total_searches", "total_crawls", "total_fetches", "avg_latency_ms", "_latency_sum", "_lock", "tool_calls")

    def __init__(self):
        self.total_searches = self.total_crawls = self.total_fetches = 0
        self.avg_latency_ms = self._latency_sum = 0.0
        self._lock = asyncio.Lock()
        self.tool_calls: dict[str, int] = {}

    async def record_search(self, latency_ms: float) -> None:
        async with self._lock:
            self.total_searches += 1
            self._latency_sum += latency_ms
            self.avg_latency_ms = self._latency_sum / self.total_searches

    async def record_crawl(self) -> None:
        async with self._lock:
            self.total_crawls += 1

    async def record_fetch(self) -> None:
        async with self._lock:
            self.total_fetches += 1

    def record_tool(self, name: str) -> None:
        self.tool_calls[name] = self.tool_calls.get(name, 0) + 1

    def to_dict(self) -> dict[str, object]:
        return {"total_searches": self.total_searches, "total_crawls": self.total_crawls, "total_fetches": self.total_fetches,
                "avg_latency_ms": round(self.avg_latency_ms, 1)}


class WebhookRegistry:
    __slots__ = ("_urls", "_max_registrations", "_secret")

    def __init__(self, max_registrations: int = _WEBHOOK_MAX_REGISTRATIONS, secret: str = ""):
        self._urls: list[str] = []
        self._max_registrations, self._secret = max_registrations, secret

    def register(self, url: str) -> str | None:
        """None on success (or already present), else the reason as text."""
        from infomesh_b200.security import SSRFError, validate_url

        try:
            validate_url(url)
        except SSRFError:
            return f"Webhook URL blocked for security: {url}"
        if url in self._urls:
            return None
        if len(self._urls) >= self._max_registrations:
            return f"Max webhooks ({self._max_registrations}) reached. Unregister one first."
        self._urls.append(url)
        return None

    def unregister(self, url: str) -> bool:
        if url in self._urls:
            self._urls.remove(url)
            return True
        return False

    @property
    def urls(self) -> list[str]:
        return list(self._urls)

    async def notify(self, event: str, payload: dict[str, object]) -> int:
        """POST ``{event, data, timestamp}`` to every hook in parallel (re-validated with DNS resolution at send time;
        HMAC-signed when a secret is configured).  Returns the number of 2xx/3xx deliveries."""
        if not self._urls:
            return 0
        import httpx

        from infomesh_b200.security import SSRFError, validate_url
        from infomesh_b200.security_ext import sign_webhook_payload

        body: dict[str, object] = {"event": event, "data": payload, "timestamp": time.time()}
        headers = {"X-InfoMesh-Signature": sign_webhook_payload(body, self._secret)} if self._secret else {}

        async def post(client: "httpx.AsyncClient", url: str) -> bool:
            try:
                validate_url(url, resolve_dns=True)
                return (await client.post(url, json=body, headers=headers)).status_code < 400
            except (SSRFError, Exception):  # noqa: BLE001
                return False

        async with httpx.AsyncClient(timeout=5.0) as client:
            return sum(await asyncio.gather(*(post(client, u) for u in self._urls)))
