"""Serving summarisation to peers over ``/infomesh/llm/1.0.0``: trust floor 0.3, 16 000-char inputs, 3 concurrent
requests per node, 5 pending + 10 s cool-down per peer, 120 s generation timeout
(reference infomesh/summarizer/peer_handler.py:30-368)."""
from __future__ import annotations

import asyncio
import time
from dataclasses import dataclass
from enum import StrEnum
from typing import Any

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

PROTOCOL_LLM = "/infomesh/llm/1.0.0"
MAX_PENDING_PER_PEER = 5
MAX_CONCURRENT_REQUESTS = 3
MAX_TEXT_LENGTH = 16_000
REQUEST_TIMEOUT_SECONDS = 120.0
PEER_COOLDOWN_SECONDS = 10.0


class RequestStatus(StrEnum):
    PENDING = "pending"
    PROCESSING = "processing"
    COMPLETED = "completed"
    REJECTED = "rejected"
    FAILED = "failed"
    TIMEOUT = "timeout"


class RejectReason(StrEnum):
    NO_LLM = "no_llm"
    RATE_LIMITED = "rate_limited"
    CAPACITY_FULL = "capacity_full"
    TEXT_TOO_LONG = "text_too_long"
    UNTRUSTED_PEER = "untrusted_peer"
    COOLDOWN = "cooldown"


@dataclass(frozen=True)
class SummarizeRequest:
    request_id: str
    requester_peer_id: str
    url: str
    title: str
    text: str
    max_tokens: int = 512
    timestamp: float = 0.0


@dataclass(frozen=True)
class SummarizeResponse:
    request_id: str
    status: RequestStatus
    summary: str | None = None
    content_hash: str | None = None
    model: str | None = None
    elapsed_ms: float = 0.0
    reject_reason: RejectReason | None = None
    detail: str = ""


class PeerSummarizationHandler:
    def __init__(self, engine, *, min_trust_score: float = 0.3):
        self._engine, self._min_trust = engine, min_trust_score
        self._active = 0
        self._last_seen: dict[str, float] = {}
        self._pending: dict[str, int] = {}
        self._last_cleanup = time.time()
        self._cleanup_interval = 3600.0
        self._served = self._rejected = 0

    active_count = property(lambda self: self._active)
    total_served = property(lambda self: self._served)
    total_rejected = property(lambda self: self._rejected)

    def _reject(self, req: SummarizeRequest, why: RejectReason, detail: str) -> SummarizeResponse:
        return SummarizeResponse(req.request_id, RequestStatus.REJECTED, reject_reason=why, detail=detail)

    def _check_rejection(self, req: SummarizeRequest, *, requester_trust: float, now: float) -> SummarizeResponse | None:
        if self._engine is None:
            return self._reject(req, RejectReason.NO_LLM, "no local LLM")
        if requester_trust < self._min_trust:
            return self._reject(req, RejectReason.UNTRUSTED_PEER,
                                f"trust score {requester_trust:.3f} below minimum {self._min_trust:.3f}")
        if len(req.text) > MAX_TEXT_LENGTH:
            return self._reject(req, RejectReason.TEXT_TOO_LONG, f"text length {len(req.text)} exceeds max {MAX_TEXT_LENGTH}")
        if self._active >= MAX_CONCURRENT_REQUESTS:
            return self._reject(req, RejectReason.CAPACITY_FULL, f"active requests {self._active}/{MAX_CONCURRENT_REQUESTS}")
        pending = self._pending.get(req.requester_peer_id, 0)
        if pending >= MAX_PENDING_PER_PEER:
            return self._reject(req, RejectReason.RATE_LIMITED, f"peer has {pending} pending requests")
        wait = PEER_COOLDOWN_SECONDS - (now - self._last_seen.get(req.requester_peer_id, 0.0))
        if wait > 0:
            return self._reject(req, RejectReason.COOLDOWN, f"cooldown: {wait:.1f}s remaining")
        return None

    def _prune_stale_peers(self, now: float) -> None:
        for pid in [p for p, ts in self._last_seen.items() if ts < now - self._cleanup_interval]:
            del self._last_seen[pid]
            self._pending.pop(pid, None)

    async def handle_request(self, request: SummarizeRequest, *, requester_trust: float = 0.5) -> SummarizeResponse:
        now = time.time()
        if now - self._last_cleanup > self._cleanup_interval:
            self._prune_stale_peers(now)
            self._last_cleanup = now
        no = self._check_rejection(request, requester_trust=requester_trust, now=now)
        if no is not None:
            self._rejected += 1
            return no
        pid = request.requester_peer_id
        self._active += 1
        self._pending[pid] = self._pending.get(pid, 0) + 1
        self._last_seen[pid] = now
        try:
            res = await asyncio.wait_for(self._engine.summarize(url=request.url, title=request.title, text=request.text,
                                                                max_tokens=request.max_tokens), timeout=REQUEST_TIMEOUT_SECONDS)
            self._served += 1
            return SummarizeResponse(request.request_id, RequestStatus.COMPLETED, res.summary, res.content_hash, res.model,
                                     res.elapsed_ms, detail="ok")
        except (TimeoutError, asyncio.TimeoutError):
            return SummarizeResponse(request.request_id, RequestStatus.TIMEOUT, detail="LLM generation timed out")
        except Exception as exc:  # noqa: BLE001
            logger.warning("peer_llm_failed", request_id=request.request_id, error=str(exc))
            return SummarizeResponse(request.request_id, RequestStatus.FAILED, detail=str(exc))
        finally:
            self._active -= 1
            self._pending[pid] = max(0, self._pending.get(pid, 1) - 1)

    async def handle_payload(self, payload: dict[str, Any], sender_peer_id: str, *, requester_trust: float = 0.5) -> dict[str, Any]:
        """Adapter for ``InfoMeshNode(llm_handler=...)``: wire dict in, wire dict out."""
        req = deserialize_request({**payload, "requester_peer_id": sender_peer_id or payload.get("requester_peer_id", "")})
        return serialize_response(await self.handle_request(req, requester_trust=requester_trust))


def serialize_request(req: SummarizeRequest) -> dict[str, Any]:
    return {"request_id": req.request_id, "requester_peer_id": req.requester_peer_id, "url": req.url, "title": req.title,
            "text": req.text, "max_tokens": req.max_tokens, "timestamp": req.timestamp}


def deserialize_request(data: dict[str, Any]) -> SummarizeRequest:
    mt = data.get("max_tokens", 512)
    return SummarizeRequest(str(data.get("request_id", "")), str(data.get("requester_peer_id", "")), str(data.get("url", "")),
                            str(data.get("title", "")), str(data.get("text", "")),
                            mt if isinstance(mt, int) and 0 < mt <= 4096 else 512, float(data.get("timestamp", 0.0) or 0.0))


def serialize_response(resp: SummarizeResponse) -> dict[str, Any]:
    return {"request_id": resp.request_id, "status": resp.status.value, "summary": resp.summary,
            "content_hash": resp.content_hash, "model": resp.model, "elapsed_ms": resp.elapsed_ms,
            "reject_reason": resp.reject_reason.value if resp.reject_reason else None, "detail": resp.detail}


def deserialize_response(data: dict[str, Any]) -> SummarizeResponse:
    rr = data.get("reject_reason")
    return SummarizeResponse(str(data.get("request_id", "")), RequestStatus(data.get("status", "failed")), data.get("summary"),
                             data.get("content_hash"), data.get("model"), float(data.get("elapsed_ms", 0.0) or 0.0),
                             RejectReason(rr) if rr else None, str(data.get("detail", "")))
