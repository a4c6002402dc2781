"""Tool implementations behind the MCP server, the HTTP API and the SDK (reference infomesh/mcp/handlers.py:66-1616).

One :class:`ToolRuntime` owns the per-process state (caches, sessions, analytics, webhooks, persistent store) and
exposes ``await runtime.call(name, arguments) -> str``.  The reference threads ~15 keyword arguments through 20 free
functions; binding them once keeps every handler to its own logic.

Search routing for ``web_search`` / ``search``:
  GPU index present and fresh  -> fused hybrid pipeline (encoder + dense + BM25 + RRF + cross-encoder rerank)
  distributed index configured -> local + DHT/peer results, cross-validated across peers
  vector store present         -> FTS5 + vector RRF hybrid
  otherwise                    -> FTS5 BM25 + 6-signal ranking
"""
from __future__ import annotations

import contextlib
import json
import time
from dataclasses import dataclass, replace
from typing import Any
from urllib.parse import urlparse

from infomesh_b200 import __version__ as SERVER_VERSION
from infomesh_b200.mcp.session import AnalyticsTracker, SessionStore, WebhookRegistry
from infomesh_b200.mcp.tools import extract_filters
from infomesh_b200.search.cache import QueryCache
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

MCP_API_VERSION = "2025.1"
_TIER_HIGH_THRESHOLD, _TIER_MID_THRESHOLD = 1000, 100
_SEARCH_ATTRIBUTION = ("\n---\nAttribution: All results sourced from their original publishers.\n"
                       "Each result includes a source URL — always cite the original source.")
_FETCH_COPYRIGHT_NOTICE = ("\n---\nCOPYRIGHT NOTICE: This content is cached by InfoMesh for search indexing purposes only.\n"
                           "The original content is owned by its respective author/publisher.\n"
                           "Always cite the original source URL when referencing this content.\n"
                           "Cache policy: content is refreshed every 7 days; may not reflect the latest version.")


class ErrorCode:
    INVALID_PARAM = "INVALID_PARAM"
    AUTH_FAILED = "AUTH_FAILED"
    SSRF_BLOCKED = "SSRF_BLOCKED"
    CRAWL_FAILED = "CRAWL_FAILED"
    FETCH_FAILED = "FETCH_FAILED"
    PAYWALL = "PAYWALL_DETECTED"
    RATE_LIMITED = "RATE_LIMITED"
    EMPTY_INDEX = "EMPTY_INDEX"
    NOT_FOUND = "NOT_FOUND"
    INTERNAL = "INTERNAL_ERROR"
    WORKER_UNAVAILABLE = "WORKER_UNAVAILABLE"


class ToolError(Exception):
    def __init__(self, code: str, message: str, hint: str = ""):
        super().__init__(message)
        self.code, self.message, self.hint = code, message, hint

    def render(self) -> str:
        return f"Error [{self.code}]: {self.message}" + (f"\nHint: {self.hint}" if self.hint else "")


def _compute_credit_tier(score: float) -> int:
    return 3 if score >= _TIER_HIGH_THRESHOLD else 2 if score >= _TIER_MID_THRESHOLD else 1


def _js(data: Any) -> str:
    return json.dumps(data, ensure_ascii=False, default=str)


def _int(args: dict[str, Any], key: str, default: int, lo: int, hi: int) -> int:
    try:
        return max(lo, min(int(args.get(key, default)), hi))
    except (TypeError, ValueError):
        return default


@dataclass
class _SearchParams:
    query: str
    original_query: str
    limit: int
    offset: int
    fmt: str
    snippet_len: int
    session_id: str | None
    filters: dict[str, Any]
    cache_key: str


class ToolRuntime:
    def __init__(self, ctx: Any, *, distributed_index: Any | None = None, p2p_node: Any | None = None, pstore: Any | None = None,
                 gpu_index: Any | None = None):
        self.ctx, self.config = ctx, ctx.config
        self.distributed_index = distributed_index if distributed_index is not None else getattr(ctx, "distributed_index", None)
        self.p2p_node = p2p_node if p2p_node is not None else getattr(ctx, "p2p_node", None)
        self.gpu_index = gpu_index if gpu_index is not None else getattr(ctx, "gpu_index", None)
        scfg = getattr(self.config, "search", None)
        self.query_cache = QueryCache(int(getattr(scfg, "cache_max_size", 1000)), float(getattr(scfg, "cache_ttl_seconds", 300.0)))
        self.sessions, self.analytics, self.webhooks = SessionStore(), AnalyticsTracker(), WebhookRegistry()
        self.pstore = pstore
        self.last_query = ""
        self._cross_encoder = None
        self._handlers = {
            "web_search": self.web_search, "fetch_page": self.fetch_page, "crawl_url": self.crawl_url, "fact_check": self.fact_check,
            "status": self.status, "search": self._legacy_search, "search_local": self._legacy_search_local,
            "network_stats": self.network_stats, "batch_search": self.batch_search, "suggest": self.suggest,
            "register_webhook": self.register_webhook, "unregister_webhook": self.unregister_webhook, "analytics": self.analytics_tool,
            "explain": self.explain, "search_history": self.search_history, "search_rag": self.search_rag,
            "extract_answer": self.extract_answer, "ping": self.ping, "credit_balance": self.credit_balance,
            "index_stats": self.index_stats, "remove_url": self.remove_url,
        }

    # ------------------------------------------------------------------ dispatch
    @property
    def tool_names(self) -> list[str]:
        return list(self._handlers)

    async def call(self, name: str, arguments: dict[str, Any] | None = None) -> str:
        fn = self._handlers.get(name)
        if fn is None:
            return ToolError(ErrorCode.NOT_FOUND, f"Unknown tool: {name}", hint=f"Available: {', '.join(list(self._handlers)[:5])}").render()
        self.analytics.record_tool(name)
        try:
            return await fn(dict(arguments or {}))
        except ToolError as err:
            return err.render()
        except Exception:  # noqa: BLE001
            logger.exception("tool_unhandled_error", tool=name)
            return f"Error [INTERNAL]: An unexpected error occurred in tool '{name}'. Please try again or report the issue."

    # ------------------------------------------------------------------ helpers
    def _authority_fn(self):
        lg = self.ctx.link_graph
        return lg.url_authority if lg else None

    def _deduct(self) -> None:
        ledger = self.ctx.ledger
        if ledger is None:
            return
        with contextlib.suppress(Exception):
            ledger.spend(ledger.search_allowance().search_cost, reason="search")

    def _truncate(self, text: str) -> str:
        cap = self.config.mcp.max_response_chars
        return text[:cap] + "\n... (truncated)" if cap > 0 and len(text) > cap else text

    def _parse_search(self, args: dict[str, Any]) -> _SearchParams:
        from infomesh_b200.search.nlp import expand_query, parse_natural_query, remove_stop_words

        query = args.get("query", "")
        if not isinstance(query, str) or not query.strip():
            raise ToolError(ErrorCode.INVALID_PARAM, "query must be a non-empty string")
        query = original = query[:1000]
        limit, offset = _int(args, "limit", 10, 1, 100), _int(args, "offset", 0, 0, 10 ** 6)
        fmt, snip = str(args.get("format", self.config.mcp.default_format)), _int(args, "snippet_length", 200, 10, 1000)
        filters = extract_filters(args)
        nq = parse_natural_query(query)           # "python docs from last week site:docs.python.org"
        if nq.date_from and "date_from" not in filters:
            filters["date_from"] = nq.date_from
        if nq.include_domains and "include_domains" not in filters:
            filters["include_domains"] = nq.include_domains
        if nq.language and "language" not in filters:
            filters["language"] = nq.language
        query = nq.cleaned_query or query
        kept = remove_stop_words(query.split())
        if kept:
            query = " ".join(kept)
        extra = expand_query(query)
        if extra:
            query = f"{query} {' '.join(extra)}"
        query = query[:1000]
        fkey = "|".join(str(x) for x in (fmt, offset, snip, filters.get("language", ""), filters.get("date_from", ""), filters.get("date_to", ""),
                                         ",".join(sorted(filters.get("include_domains", []))), ",".join(sorted(filters.get("exclude_domains", [])))))
        sid = args.get("session_id")
        return _SearchParams(query, original, limit, offset, fmt, snip, sid if isinstance(sid, str) else None, filters,
                             QueryCache.make_key(query, limit, f=fkey))

    def _post_search(self, text: str, p: _SearchParams) -> str:
        from infomesh_b200.search.nlp import did_you_mean

        ledger, store = self.ctx.ledger, self.ctx.store
        if p.fmt == "json" and ledger is not None:
            with contextlib.suppress(Exception):
                data = json.loads(text)
                al = ledger.search_allowance()
                data["quota"] = {"credit_balance": round(ledger.balance(), 2), "state": al.state.value, "search_cost": round(al.search_cost, 3)}
                data["api_version"] = MCP_API_VERSION
                text = _js(data)
        if text.startswith("No results found"):
            with contextlib.suppress(Exception):
                if store.get_stats().get("document_count", 0) == 0:
                    text += ("\n\nYour index is empty. Try crawling some pages first:\n"
                             "  crawl_url(url='https://docs.python.org/3/', depth=1)")
            with contextlib.suppress(Exception):
                sugg = did_you_mean(p.original_query, store.suggest(p.original_query[:20], limit=50))
                if sugg:
                    text += f'\n\nDid you mean: "{sugg[0] if isinstance(sugg, list) else sugg}"?'
        elif p.fmt != "json" and self.config.mcp.show_attribution:
            text += _SEARCH_ATTRIBUTION
        return self._truncate(text)

    def _reranker(self):
        if self._cross_encoder is None and self.gpu_index is not None and self.gpu_index.reranker is not None:
            from infomesh_b200.search.reranker import CrossEncoderReranker

            self._cross_encoder = CrossEncoderReranker(self.gpu_index.reranker, device=str(self.gpu_index.device))
        return self._cross_encoder

    async def _rerank(self, query: str, results: list[Any], enabled: bool) -> list[Any]:
        if not enabled or not results:
            return results
        ce = self._reranker()
        if ce is not None:
            import asyncio

            from infomesh_b200.search.reranker import rerank_with_cross_encoder

            return await asyncio.to_thread(rerank_with_cross_encoder, query, results, ce)
        if self.ctx.llm_backend is not None:
            from infomesh_b200.search.reranker import rerank_with_llm

            return await rerank_with_llm(query, results, self.ctx.llm_backend)
        return results

    # ------------------------------------------------------------------ search core
    async def _search(self, args: dict[str, Any], *, network: bool, rerank: bool = True) -> str:
        from infomesh_b200.search import formatter as F
        from infomesh_b200.search import query as Q

        p = self._parse_search(args)
        ck = f"{p.cache_key}:{int(network)}:{int(rerank)}"
        hit = self.query_cache.get(ck)
        if hit is not None:
            return hit
        t0 = time.monotonic()
        self._deduct()
        store, vstore, auth = self.ctx.store, self.ctx.vector_store, self._authority_fn()
        as_json = p.fmt == "json"
        gi = self.gpu_index
        if gi is not None and gi.engine is not None and not p.filters and p.offset == 0 and not (network and self.distributed_index is not None):
            import asyncio

            hits = await asyncio.to_thread(gi.search, p.query, p.limit)
            ms = (time.monotonic() - t0) * 1000
            text = (_js({"total": len(hits), "elapsed_ms": round(ms, 1), "source": "gpu_hybrid", "results": hits}) if as_json
                    else _format_gpu_hits(hits, ms, p.snippet_len))
        elif network and self.distributed_index is not None:
            nsf = getattr(self.p2p_node, "search_network", None)
            dist = await Q.search_distributed(store, self.distributed_index, p.query, limit=p.limit, authority_fn=auth, vector_store=vstore,
                                              network_search_fn=nsf)
            if dist.remote_count > 0:
                self._cross_validate(p.query, dist)
            dist = replace(dist, results=await self._rerank(p.query, dist.results, rerank))
            text = (F.format_distributed_results_json if as_json else F.format_distributed_results)(dist, max_snippet=p.snippet_len)
        elif network and vstore is not None:
            hyb = Q.search_hybrid(store, vstore, p.query, limit=p.limit, authority_fn=auth)
            text = (F.format_hybrid_results_json if as_json else F.format_hybrid_results)(hyb, max_snippet=p.snippet_len)
        else:
            res = Q.search_local(store, p.query, limit=p.limit, offset=p.offset, authority_fn=auth, **p.filters)
            res = replace(res, results=await self._rerank(p.query, res.results, rerank))
            text = (F.format_fts_results_json if as_json else F.format_fts_results)(res, max_snippet=p.snippet_len)
        ms = (time.monotonic() - t0) * 1000
        await self.analytics.record_search(ms)
        if self.pstore is not None:
            with contextlib.suppress(Exception):
                self.pstore.record_search(ms)
        fb = getattr(self.ctx, "feedback_store", None)
        if fb is not None and self.config.search.feedback_tracking:
            with contextlib.suppress(Exception):
                if self.last_query and self.last_query != p.original_query and fb.is_reformulation(self.last_query):
                    fb.record_reformulation(self.last_query)
        self.last_query = p.original_query
        text = self._post_search(text, p)
        if p.session_id:
            s = self.sessions.get_or_create(p.session_id)
            s.last_query, s.last_results, s.updated_at = p.query, text[:2000], time.time()
        self.query_cache.put(ck, text)
        return text

    def _cross_validate(self, query: str, dist: Any) -> None:
        from infomesh_b200.search.cross_validate import PeerResult, cross_validate_results

        by_peer: dict[str, list[PeerResult]] = {}
        for r in dist.results:
            pid = getattr(r, "peer_id", None) or "local"
            by_peer.setdefault(pid, []).append(PeerResult(pid, r.url, r.title, r.snippet, r.combined_score))
        rep = cross_validate_results(query, by_peer)
        if rep.suspicious_count or rep.fabricated_count:
            logger.warning("cross_validate_suspicious", query=query[:60], suspicious=rep.suspicious_count, fabricated=rep.fabricated_count)

    async def _legacy_search(self, args):
        return await self._search(args, network=True)

    async def _legacy_search_local(self, args):
        return await self._search(args, network=False)

    # ------------------------------------------------------------------ the five tools
    async def web_search(self, args: dict[str, Any]) -> str:
        query = args.get("query", "")
        if not isinstance(query, str) or not query.strip():
            raise ToolError(ErrorCode.INVALID_PARAM, "query must be a non-empty string")
        top_k = _int(args, "top_k", 5, 1, 100)
        filters = extract_filters(args)
        base = {"query": query, "limit": top_k, **filters}
        if args.get("explain"):
            return await self.explain({**base, "format": "json"})
        if args.get("chunk_size") is not None:
            return await self.search_rag({**base, "chunk_size": _int(args, "chunk_size", 500, 50, 8000)})
        if args.get("answer_mode", "snippets") in ("summary", "structured"):
            return await self.extract_answer(base)
        text = await self._search({**base, "format": "json"}, network=not args.get("local_only", False), rerank=bool(args.get("rerank", True)))
        if args.get("fetch_full_content"):
            with contextlib.suppress(Exception):
                data = json.loads(text)
                for r in data.get("results", []):
                    doc = self.ctx.store.get_document_by_url(r.get("url", ""))
                    if doc is not None:
                        r["full_text"] = doc.text[:10000]
                text = _js(data)
        return text

    async def fetch_page(self, args: dict[str, Any]) -> str:
        from infomesh_b200.search.formatter import format_fetch_result
        from infomesh_b200.security import SSRFError, validate_url
        from infomesh_b200.services import fetch_page_async

        url, fmt = args.get("url", ""), args.get("format", "text")
        if not url or not isinstance(url, str):
            raise ToolError(ErrorCode.INVALID_PARAM, "url must be a non-empty string")
        try:
            validate_url(url)
        except SSRFError as exc:
            raise ToolError(ErrorCode.SSRF_BLOCKED, f"URL blocked for security: {exc}") from None
        cfg = self.config
        ttl = cfg.storage.cache_ttl_days * 86400
        cached = self.ctx.store.get_document_by_url(url)
        if cached is None and self.ctx.worker is None:
            raise ToolError(ErrorCode.WORKER_UNAVAILABLE, "fetch_page requires a crawler worker", hint="Start the node with 'infomesh start' first.")
        fp = await fetch_page_async(url, store=self.ctx.store, worker=self.ctx.worker, vector_store=self.ctx.vector_store,
                                    max_size_bytes=cfg.index.max_doc_size_kb * 1024, cache_ttl_seconds=ttl)
        await self.analytics.record_fetch()
        if self.pstore is not None:
            with contextlib.suppress(Exception):
                self.pstore.record_fetch()
        fb = getattr(self.ctx, "feedback_store", None)
        if fb is not None and self.last_query and cfg.search.feedback_tracking:
            with contextlib.suppress(Exception):
                fb.record_fetch(self.last_query, url, 0)
        if not fp.success:
            if fp.is_paywall:
                raise ToolError(ErrorCode.PAYWALL, f"Paywall detected for {url}", hint="This page requires a subscription.")
            raise ToolError(ErrorCode.FETCH_FAILED, f"Failed to fetch {url}: content unavailable", hint="The page may be down. Try again later.")
        if fmt == "json":
            return _js({"url": fp.url, "title": fp.title, "domain": urlparse(fp.url).netloc, "text": fp.text, "is_cached": fp.is_cached,
                        "crawled_at": fp.crawled_at, "is_paywall": fp.is_paywall, "api_version": MCP_API_VERSION})
        text = format_fetch_result(title=fp.title, url=fp.url, text=fp.text, is_cached=fp.is_cached, crawled_at=fp.crawled_at,
                                   cache_ttl=ttl, is_paywall=fp.is_paywall)
        return self._truncate(text + (_FETCH_COPYRIGHT_NOTICE if cfg.mcp.show_copyright else ""))

    async def crawl_url(self, args: dict[str, Any]) -> str:
        from infomesh_b200.security import SSRFError, validate_url
        from infomesh_b200.services import crawl_and_index

        url = args.get("url", "")
        if not url or not isinstance(url, str):
            raise ToolError(ErrorCode.INVALID_PARAM, "url must be a non-empty string")
        worker = self.ctx.worker
        if worker is None:
            raise ToolError(ErrorCode.WORKER_UNAVAILABLE, "crawl_url requires a crawler worker", hint="Start the node with 'infomesh start' first.")
        try:
            validate_url(url)
        except SSRFError as exc:
            raise ToolError(ErrorCode.SSRF_BLOCKED, f"URL blocked for security: {exc}") from None
        depth = _int(args, "depth", 0, 0, 10)
        if self.config.crawl.max_depth > 0:
            depth = min(depth, self.config.crawl.max_depth)
        if isinstance(args.get("webhook_url"), str):
            self.webhooks.register(args["webhook_url"])
        if depth > 0 and hasattr(worker, "set_scope"):
            import asyncio

            scoped = worker.set_scope(url)
            if asyncio.iscoroutine(scoped):
                await scoped
        ci = await crawl_and_index(url, worker=worker, store=self.ctx.store, vector_store=self.ctx.vector_store, p2p_node=self.p2p_node,
                                   distributed_index=self.distributed_index, link_graph=self.ctx.link_graph, depth=depth,
                                   force=bool(args.get("force", False)))
        await self.analytics.record_crawl()
        if self.pstore is not None:
            with contextlib.suppress(Exception):
                self.pstore.record_crawl()
        if not ci.success:
            raise ToolError(ErrorCode.CRAWL_FAILED, f"Crawl failed for {url}: {ci.error}", hint="Check if the URL is reachable.")
        if self.gpu_index is not None:
            self.gpu_index.note_added()
        self.query_cache.clear()
        await self.webhooks.notify("crawl_completed", {"url": url, "title": ci.title, "text_length": ci.text_length,
                                                       "links_discovered": ci.links_discovered, "elapsed_ms": round(ci.elapsed_ms, 0)})
        return (f"Crawled successfully: {url}\nTitle: {ci.title}\nText length: {ci.text_length} chars\n"
                f"Links discovered: {ci.links_discovered}\nElapsed: {ci.elapsed_ms:.0f}ms")

    async def fact_check(self, args: dict[str, Any]) -> str:
        from infomesh_b200.data_quality import cross_reference_results
        from infomesh_b200.search.query import search_local

        claim = args.get("claim", "")
        if not claim or not isinstance(claim, str):
            raise ToolError(ErrorCode.INVALID_PARAM, "claim is required")
        limit = _int(args, "top_k", _int(args, "limit", 10, 1, 50), 1, 50)
        res = search_local(self.ctx.store, claim, limit=limit, authority_fn=self._authority_fn(), **extract_filters(args))
        fc = cross_reference_results(claim, res.results)
        fb = getattr(self.ctx, "feedback_store", None)
        if fb is not None and self.config.search.feedback_tracking:
            with contextlib.suppress(Exception):
                for u in fc.sources[:3]:
                    fb.record_citation(claim, u)
        return _js({"api_version": MCP_API_VERSION, "claim": claim, "verdict": fc.verdict, "confidence": round(fc.confidence, 3),
                    "supporting": fc.supporting_sources, "contradicting": fc.contradicting_sources, "sources_checked": len(res.results),
                    "sources": fc.sources})

    def _status_data(self) -> dict[str, object]:
        ctx = self.ctx
        data: dict[str, object] = {"api_version": MCP_API_VERSION, "phase": "4 (Production)",
                                   "documents_indexed": ctx.store.get_stats().get("document_count", 0),
                                   "pending_crawl_urls": ctx.scheduler.pending_count if ctx.scheduler else 0,
                                   "ranking": "BM25 + freshness + trust + authority", "analytics": self.analytics.to_dict()}
        if ctx.vector_store is not None:
            vs = ctx.vector_store.get_stats()
            data["vector"] = {"documents": vs.get("document_count", 0), "model": vs.get("model", "unknown")}
        else:
            data["vector"] = {"enabled": False}
        if ctx.link_graph:
            lg = ctx.link_graph.get_stats()
            data["link_graph"] = {"links": lg.get("link_count", 0), "domains_scored": lg.get("domain_count", 0)}
        else:
            data["link_graph"] = {"enabled": False}
        if ctx.ledger is not None:
            data["credits"] = self._credit_data()
        data["p2p"] = self._p2p_status()
        data["gpu"] = self.gpu_index.stats() if self.gpu_index is not None else {"enabled": False}
        return data

    def _credit_data(self) -> dict[str, object]:
        ledger = self.ctx.ledger
        al = ledger.search_allowance()
        cr: dict[str, object] = {"balance": round(ledger.balance(), 2), "state": al.state.value, "search_cost": round(al.search_cost, 3),
                                 "tier": _compute_credit_tier(float(ledger.contribution_score()))}
        if al.state.value == "grace":
            cr["grace_remaining_hours"] = round(al.grace_remaining_hours or 0.0, 1)
        elif al.state.value == "debt":
            cr["debt_amount"] = round(al.debt_amount, 2)
        mgr = getattr(self.ctx, "credit_sync_manager", None)
        if mgr is not None:
            with contextlib.suppress(Exception):
                agg = mgr.aggregated_stats()
                if agg.node_count > 1:
                    cr["network"] = {"total_earned": round(agg.total_earned, 2), "total_spent": round(agg.total_spent, 2),
                                     "balance": round(agg.balance, 2), "contribution_score": round(agg.contribution_score, 2),
                                     "node_count": agg.node_count}
        return cr

    def _p2p_status(self) -> dict[str, object]:
        node = self.p2p_node
        if node is None:
            return {"peers": 0, "mode": "local"}
        try:
            out: dict[str, object] = {"peers": len(node.connected_peers)}
            if self.distributed_index is not None:
                di = self.distributed_index.stats
                out["dht"] = {"published": di.documents_published, "keywords": di.keywords_published, "queries": di.queries_performed}
            return out
        except Exception:  # noqa: BLE001
            return {"error": "status unavailable"}

    async def status(self, args: dict[str, Any]) -> str:
        data = self._status_data()
        data.update(status="ok", server="infomesh", version=SERVER_VERSION)
        with contextlib.suppress(Exception):
            data["top_domains"] = [{"domain": d, "count": c} for d, c in self.ctx.store.get_top_domains(limit=5)]
        return _js(data)

    # ------------------------------------------------------------------ legacy tools
    async def network_stats(self, args: dict[str, Any]) -> str:
        data = self._status_data()
        if args.get("format") == "json":
            return _js(data)
        lines = [f"InfoMesh Node Status (API {MCP_API_VERSION})", f"  Documents indexed: {data['documents_indexed']}",
                 f"  Pending crawl URLs: {data['pending_crawl_urls']}", f"  Ranking: {data['ranking']}",
                 f"  Peers: {data['p2p'].get('peers', 0) if isinstance(data['p2p'], dict) else 0}"]
        if "credits" in data:
            cr = data["credits"]
            lines.append(f"  Credits: {cr['balance']} ({cr['state']}, tier {cr['tier']})")
        return "\n".join(lines)

    async def batch_search(self, args: dict[str, Any]) -> str:
        """Up to 10 queries (extra ones are ignored, as in reference mcp/handlers.py:906-977).  ``format="text"`` (default) renders
        each query with the standard result formatter; ``format="json"`` returns ``{"api_version", "batch_results": [...]}``."""
        from infomesh_b200.search.formatter import format_fts_results, format_fts_results_json
        from infomesh_b200.search.query import search_local

        queries = args.get("queries", [])
        if not isinstance(queries, list) or not queries:
            raise ToolError(ErrorCode.INVALID_PARAM, "queries must be a non-empty array")
        if len(queries) > 50:
            raise ToolError(ErrorCode.INVALID_PARAM, "Batch search exceeds maximum queries (50)")
        queries, limit, fmt = queries[:10], _int(args, "limit", 5, 1, 50), args.get("format", "text")
        gi = self.gpu_index
        if gi is not None and gi.engine is not None and fmt == "json":          # one device pass for the whole batch
            import asyncio

            rows = await asyncio.to_thread(gi.search_many, [str(q) for q in queries], limit)
            batch = [{"query": q, "source": "gpu_hybrid", "total": len(r), "results": r} for q, r in zip(queries, rows)]
            return _js({"api_version": MCP_API_VERSION, "batch_results": batch, "results": batch})
        batch, parts = [], []
        for i, q in enumerate(queries, 1):
            if not isinstance(q, str) or not q.strip():
                batch.append({"query": str(q), "error": "invalid"})
                parts.append(f"--- Query {i}: (invalid) ---\n")
                continue
            self._deduct()
            t0 = time.monotonic()
            res = search_local(self.ctx.store, q, limit=limit, authority_fn=self._authority_fn())
            await self.analytics.record_search((time.monotonic() - t0) * 1000)
            if fmt == "json":
                batch.append({**json.loads(format_fts_results_json(res)), "query": q})
            else:
                parts.append(f"--- Query {i}: {q} ---\n{format_fts_results(res)}\n")
        if fmt == "json":
            return _js({"api_version": MCP_API_VERSION, "batch_results": batch, "results": batch})
        return "\n".join(parts)

    async def suggest(self, args: dict[str, Any]) -> str:
        prefix = args.get("prefix", "")
        if not isinstance(prefix, str) or not prefix.strip():
            raise ToolError(ErrorCode.INVALID_PARAM, "prefix must be a non-empty string")
        return _js({"prefix": prefix, "suggestions": self.ctx.store.suggest(prefix[:100], limit=_int(args, "limit", 10, 1, 50))})

    async def register_webhook(self, args: dict[str, Any]) -> str:
        url = args.get("url", "")
        if not url or not isinstance(url, str):
            raise ToolError(ErrorCode.INVALID_PARAM, "url must be a non-empty string")
        err = self.webhooks.register(url)
        if err:
            raise ToolError(ErrorCode.SSRF_BLOCKED if "blocked" in err else ErrorCode.RATE_LIMITED, err)
        if self.pstore is not None:
            self.pstore.register_webhook(url)
        return f"Webhook registered: {url}"

    async def unregister_webhook(self, args: dict[str, Any]) -> str:
        url = str(args.get("url", ""))
        ok = self.webhooks.unregister(url)
        if self.pstore is not None:
            ok = self.pstore.unregister_webhook(url) or ok
        return f"Webhook {'removed' if ok else 'not found'}: {url}"

    async def analytics_tool(self, args: dict[str, Any]) -> str:
        data = dict(self.analytics.to_dict())
        if self.pstore is not None:
            data["persistent"] = self.pstore.get_analytics()
        data["tools"] = dict(self.analytics.tool_calls)
        data["cache"] = dict(self.query_cache.stats.__dict__)
        return _js(data)

    async def explain(self, args: dict[str, Any]) -> str:
        from infomesh_b200.search.explain import explain_query
        from infomesh_b200.search.query import _sanitize_fts_query, search_local

        query = args.get("query", "")
        if not isinstance(query, str) or not query.strip():
            raise ToolError(ErrorCode.INVALID_PARAM, "query must be a non-empty string")
        res = search_local(self.ctx.store, query, limit=_int(args, "limit", 5, 1, 50), authority_fn=self._authority_fn(), **extract_filters(args))
        ex = explain_query(query, _sanitize_fts_query(query), res.results, res.elapsed_ms)
        if args.get("format", "json") == "text":
            lines = [f"Query: {ex.query}", f"Results: {ex.total_results}", f"Elapsed: {ex.elapsed_ms:.1f}ms", ""]
            for e in ex.results:
                lines += [f"  {e.url}", f"    Score: {e.combined_score:.4f}", *(f"    {k}: {v:.4f}" for k, v in e.weighted.items()), ""]
            return "\n".join(lines)
        return _js({"api_version": MCP_API_VERSION, **ex.to_dict()})

    async def search_history(self, args: dict[str, Any]) -> str:
        if self.pstore is None:
            return _js({"history": []})
        if args.get("clear"):
            return _js({"cleared": self.pstore.clear_history()})
        return _js({"history": self.pstore.get_history(limit=_int(args, "limit", 20, 1, 200))})

    async def search_rag(self, args: dict[str, Any]) -> str:
        from infomesh_b200.search.query import search_local
        from infomesh_b200.search.rag import format_rag_output

        query = args.get("query", "")
        if not isinstance(query, str) or not query.strip():
            raise ToolError(ErrorCode.INVALID_PARAM, "query must be a non-empty string")
        self._deduct()
        t0 = time.monotonic()
        res = search_local(self.ctx.store, query, limit=_int(args, "limit", 5, 1, 50), authority_fn=self._authority_fn(), **extract_filters(args))
        await self.analytics.record_search((time.monotonic() - t0) * 1000)
        out = format_rag_output(query, res.results, chunk_size=_int(args, "chunk_size", 500, 50, 8000), max_chunks=_int(args, "max_chunks", 10, 1, 50))
        data = out.to_dict()
        # key names of the reference's tool output (mcp/handlers.py search_rag) next to the richer chunk records
        data["context_chunks"] = [{"url": c["url"], "title": c["title"], "text": c["text"], "relevance_score": round(float(c["score"]), 4)}
                                  for c in data.get("chunks", [])]
        return _js({"api_version": MCP_API_VERSION, **data})

    async def extract_answer(self, args: dict[str, Any]) -> str:
        from infomesh_b200.search.query import search_local
        from infomesh_b200.search.rag import extract_answers

        query = args.get("query", "")
        if not isinstance(query, str) or not query.strip():
            raise ToolError(ErrorCode.INVALID_PARAM, "query must be a non-empty string")
        self._deduct()
        res = search_local(self.ctx.store, query, limit=_int(args, "limit", 5, 1, 50), authority_fn=self._authority_fn(), **extract_filters(args))
        answers = extract_answers(query, res.results)
        return _js({"api_version": MCP_API_VERSION, "query": query,
                    "answers": [{"answer": a.answer, "source_url": a.source_url, "source_title": a.source_title, "confidence": a.confidence}
                                for a in answers],
                    "sources": [{"url": r.url, "title": r.title, "score": round(r.combined_score, 4)} for r in res.results]})

    async def ping(self, args: dict[str, Any]) -> str:
        return _js({"status": "ok", "server": "infomesh", "version": SERVER_VERSION, "api_version": MCP_API_VERSION, "timestamp": time.time()})

    async def credit_balance(self, args: dict[str, Any]) -> str:
        if self.ctx.ledger is None:
            data: dict[str, object] = {"balance": 0, "state": "normal", "search_cost": 0.1, "note": "Credit ledger not active"}
        else:
            data = self._credit_data()
            st = self.ctx.ledger.stats()
            data.update(total_earned=round(st.total_earned, 2), total_spent=round(st.total_spent, 2), contribution_score=round(st.contribution_score, 2))
        if args.get("format", "json") == "text":
            return "\n".join(["Credit Balance", "==============", f"Balance: {data.get('balance', 0)}", f"State: {data.get('state', 'n/a')}",
                              f"Search cost: {data.get('search_cost', 0)}"])
        return _js(data)

    async def index_stats(self, args: dict[str, Any]) -> str:
        store = self.ctx.store
        data: dict[str, object] = {"document_count": store.get_stats().get("document_count", 0)}
        with contextlib.suppress(Exception):
            data["top_domains"] = [{"domain": d, "count": c} for d, c in store.get_top_domains(limit=10)]
        if self.ctx.vector_store is not None:
            vs = self.ctx.vector_store.get_stats()
            data["vector"] = {**vs, "document_count": vs.get("document_count", 0), "model": vs.get("model", "unknown")}
        if self.gpu_index is not None:
            data["gpu"] = self.gpu_index.stats()
        if args.get("format", "json") == "text":
            lines = ["Index Statistics", "================", f"Documents: {data['document_count']}"]
            if data.get("top_domains"):
                lines.append("Top domains:")
                lines += [f"  {d['domain']}: {d['count']}" for d in data["top_domains"][:5]]      # type: ignore[index]
            return "\n".join(lines)
        return _js(data)

    async def remove_url(self, args: dict[str, Any]) -> str:
        url = args.get("url", "")
        if not url or not isinstance(url, str):
            raise ToolError(ErrorCode.INVALID_PARAM, "url must be a non-empty string")
        doc = self.ctx.store.get_document_by_url(url)
        if doc is None:
            raise ToolError(ErrorCode.NOT_FOUND, f"URL not in index: {url}")
        self.ctx.store.delete_document(doc.doc_id)
        if self.ctx.vector_store is not None:
            with contextlib.suppress(Exception):
                self.ctx.vector_store.delete_document(doc.doc_id)
        if self.gpu_index is not None:
            self.gpu_index.mark_deleted(doc.doc_id)
        return f"Removed from index: {url}"


def _format_gpu_hits(hits: list[dict[str, object]], elapsed_ms: float, max_snippet: int) -> str:
    if not hits:
        return "No results found."
    lines = [f"Found {len(hits)} results ({elapsed_ms:.0f}ms, gpu hybrid):", ""]
    for i, h in enumerate(hits, 1):
        lines += [f"{i}. {h['title']}", f"   {h['url']}", f"   rerank score: {float(h['score']):.3f}", f"   {str(h['snippet'])[:max_snippet]}", ""]
    return "\n".join(lines)


# ---------------------------------------------------------------------------------------------------------------------
# Functional entry points.  The reference exposes one free function per tool that takes its collaborators as keyword
# arguments (infomesh/mcp/handlers.py:218-1616); embedders that call those directly keep working: each adapter wraps the
# collaborators in a throw-away context, runs the same ToolRuntime method as the MCP server, and returns MCP text content.
# ---------------------------------------------------------------------------------------------------------------------
_CTX_FIELDS = ("store", "vector_store", "link_graph", "ledger", "llm_backend", "worker", "scheduler", "feedback_store",
               "credit_sync_manager", "distributed_index", "p2p_node", "key_pair", "dedup", "gpu_index")


def _text_content(text: str) -> list[Any]:
    try:
        from mcp.types import TextContent

        return [TextContent(type="text", text=text)]
    except ImportError:                       # the MCP SDK is optional; the shape stays the same
        from types import SimpleNamespace

        return [SimpleNamespace(type="text", text=text)]


def _adhoc_runtime(*, config: Any | None = None, query_cache: QueryCache | None = None, sessions: SessionStore | None = None,
                   analytics: AnalyticsTracker | None = None, webhooks: WebhookRegistry | None = None, last_search_query: str = "",
                   **deps: Any) -> "ToolRuntime":
    from types import SimpleNamespace

    from infomesh_b200.config import Config

    ctx = SimpleNamespace(config=config or Config(), **{f: deps.get(f) for f in _CTX_FIELDS})
    rt = ToolRuntime(ctx)
    rt.query_cache = query_cache if query_cache is not None else rt.query_cache
    rt.sessions = sessions if sessions is not None else rt.sessions
    rt.analytics = analytics if analytics is not None else rt.analytics
    rt.webhooks = webhooks if webhooks is not None else rt.webhooks
    rt.last_query = last_search_query
    return rt


async def _run_tool(tool: str, arguments: dict[str, Any] | None, **deps: Any) -> list[Any]:
    return _text_content(await _adhoc_runtime(**deps).call(tool, dict(arguments or {})))


def _run_tool_sync(tool: str, arguments: dict[str, Any] | None, **deps: Any) -> list[Any]:
    import asyncio

    coro = _run_tool(tool, arguments, **deps)
    try:
        asyncio.get_running_loop()
    except RuntimeError:
        return asyncio.run(coro)
    import concurrent.futures                 # called from inside a loop: finish on a helper thread

    with concurrent.futures.ThreadPoolExecutor(max_workers=1) as pool:
        return pool.submit(asyncio.run, coro).result()


def deduct_search_cost(ledger: Any) -> None:
    """Charge one search to the ledger; accounting problems never fail the search."""
    if ledger is None:
        return
    try:
        ledger.spend(ledger.search_allowance().search_cost, reason="search")
    except Exception:  # noqa: BLE001
        logger.debug("search_cost_deduction_failed")


async def handle_search(name: str, arguments: dict[str, Any], **deps: Any) -> list[Any]:
    """``name``: ``"search"`` (network + local) or ``"search_local"``."""
    return await _run_tool(name if name in ("search", "search_local") else "search", arguments, **deps)


async def handle_web_search(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return await _run_tool("web_search", arguments, **deps)


async def handle_fetch(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return await _run_tool("fetch_page", arguments, **deps)


async def handle_crawl(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return await _run_tool("crawl_url", arguments, **deps)


async def handle_batch(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return await _run_tool("batch_search", arguments, **deps)


async def handle_explain(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return await _run_tool("explain", arguments, **deps)


async def handle_search_rag(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return await _run_tool("search_rag", arguments, **deps)


async def handle_extract_answer(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return await _run_tool("extract_answer", arguments, **deps)


async def handle_fact_check(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return await _run_tool("fact_check", arguments, **deps)


def handle_stats(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return _run_tool_sync("network_stats", arguments, **deps)


def handle_status(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return _run_tool_sync("status", arguments, **deps)


def handle_suggest(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return _run_tool_sync("suggest", arguments, **deps)


def handle_ping() -> list[Any]:
    return _text_content(_js({"status": "ok", "server": "infomesh", "version": SERVER_VERSION, "api_version": MCP_API_VERSION}))


def handle_credit_balance(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return _run_tool_sync("credit_balance", arguments, **deps)


def handle_index_stats(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return _run_tool_sync("index_stats", arguments, **deps)


def handle_remove_url(arguments: dict[str, Any], **deps: Any) -> list[Any]:
    return _run_tool_sync("remove_url", arguments, **deps)
