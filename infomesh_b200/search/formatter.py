"""Text and JSON renderers shared by CLI / MCP / dashboard.  Output shapes are part of the public API and follow
reference infomesh/search/formatter.py:22-284 field for field."""
from __future__ import annotations

import json
import time
from typing import Any
from urllib.parse import urlparse

from infomesh_b200.data_quality import compute_freshness_indicator
from infomesh_b200.index.ranking import RankedResult
from infomesh_b200.search.merge import MergedResult


def _host(url: str) -> str:
    try:
        return urlparse(url).netloc
    except ValueError:
        return ""


def _ranked_to_dict(r: RankedResult, *, max_snippet: int = 200) -> dict[str, object]:
    d: dict[str, object] = {
        "url": r.url, "title": r.title, "domain": _host(r.url), "snippet": r.snippet[:max_snippet],
        "score": round(r.combined_score, 4),
        "scores": {"bm25": round(r.bm25_score, 4), "freshness": round(r.freshness_score, 4),
                   "trust": round(r.trust_score, 4), "authority": round(r.authority_score, 4),
                   "title_match": round(r.title_match_score, 4), "url_path": round(r.url_path_score, 4)},
        "crawled_at": r.crawled_at, "peer_id": r.peer_id,
    }
    if r.crawled_at:
        fi = compute_freshness_indicator(r.crawled_at)
        d["freshness_grade"] = fi.freshness_grade
        d["freshness_label"] = fi.age_label
    return d


def _merged_to_dict(r: MergedResult, *, max_snippet: int = 200) -> dict[str, object]:
    scores: dict[str, float] = {}
    if r.fts_score is not None:
        scores["bm25"] = round(r.fts_score, 4)
    if r.vector_score is not None:
        scores["vector"] = round(r.vector_score, 4)
    scores["rrf"] = round(r.combined_score, 4)
    return {"url": r.url, "title": r.title, "domain": _host(r.url), "snippet": r.snippet[:max_snippet],
            "source": r.source, "scores": scores}


def _envelope(result: Any, rows: list[dict[str, object]], **extra: object) -> str:
    data = {"total": result.total, "elapsed_ms": round(result.elapsed_ms, 1), "source": result.source, **extra,
            "results": rows}
    return json.dumps(data, ensure_ascii=False)


def format_fts_results_json(result: Any, *, max_snippet: int = 200) -> str:
    return _envelope(result, [_ranked_to_dict(r, max_snippet=max_snippet) for r in result.results])


def format_hybrid_results_json(hybrid: Any, *, max_snippet: int = 200) -> str:
    return _envelope(hybrid, [_merged_to_dict(r, max_snippet=max_snippet) for r in hybrid.results])


def format_distributed_results_json(result: Any, *, max_snippet: int = 200) -> str:
    return _envelope(result, [_ranked_to_dict(r, max_snippet=max_snippet) for r in result.results],
                     local_count=result.local_count, remote_count=result.remote_count)


def _freshness_label(crawled_at: float | None) -> str:
    return f", {compute_freshness_indicator(crawled_at).age_label}" if crawled_at else ""


def _format_ranked(idx: int, r: RankedResult, *, max_snippet: int = 200) -> str:
    peer = f"    Peer: {r.peer_id}\n" if r.peer_id else ""
    return (f"[{idx}] {r.title}\n    Source: {r.url}\n    Domain: {_host(r.url)}\n{peer}"
            f"    Score: {r.combined_score:.4f} (BM25={r.bm25_score:.3f}, fresh={r.freshness_score:.3f}, "
            f"trust={r.trust_score:.3f}, auth={r.authority_score:.3f}{_freshness_label(r.crawled_at)})\n"
            f"    {r.snippet[:max_snippet]}\n")


def _format_merged(idx: int, r: MergedResult, *, max_snippet: int = 200) -> str:
    parts = []
    if r.fts_score is not None:
        parts.append(f"BM25={r.fts_score:.3f}")
    if r.vector_score is not None:
        parts.append(f"sim={r.vector_score:.3f}")
    return (f"[{idx}] {r.title} [{r.source}]\n    Source: {r.url}\n    Domain: {_host(r.url)}\n"
            f"    Score: {', '.join(parts) if parts else 'N/A'} (RRF={r.combined_score:.4f})\n"
            f"    {r.snippet[:max_snippet]}\n")


def format_fts_results(result: Any, *, max_snippet: int = 200) -> str:
    if not result.results:
        return "No results found."
    head = f"Found {result.total} results ({result.elapsed_ms:.0f}ms):\n"
    return "\n".join([head, *(_format_ranked(i, r, max_snippet=max_snippet) for i, r in enumerate(result.results, 1))])


def format_hybrid_results(hybrid: Any, *, max_snippet: int = 200) -> str:
    if not hybrid.results:
        return "No results found."
    head = f"Found {hybrid.total} results ({hybrid.elapsed_ms:.0f}ms, {hybrid.source}):\n"
    return "\n".join([head, *(_format_merged(i, r, max_snippet=max_snippet) for i, r in enumerate(hybrid.results, 1))])


def format_distributed_results(result: Any, *, max_snippet: int = 200) -> str:
    if not result.results:
        return "No results found."
    head = (f"Found {result.total} results ({result.elapsed_ms:.0f}ms, {result.source})\n"
            f"  Local: {result.local_count}, Remote: {result.remote_count}\n")
    return "\n".join([head, *(_format_ranked(i, r, max_snippet=max_snippet) for i, r in enumerate(result.results, 1))])


def format_fetch_result(*, title: str, url: str, text: str, is_cached: bool, crawled_at: float = 0.0,
                        cache_ttl: float = 604_800, is_paywall: bool = False) -> str:
    """Metadata header + page text for the ``fetch_page`` tool."""
    head = [f"# {title}", f"Source: {url}", f"Domain: {_host(url)}"]
    if is_cached:
        age = time.time() - crawled_at
        head += ["is_cached: true", f"cache_age: {age / 86400:.1f} days", f"crawl_timestamp: {crawled_at:.0f}"]
        if age > cache_ttl:
            head.append("stale_warning: true (cached content older than TTL)")
    else:
        head += ["is_cached: false", "cache_age: 0 days (freshly crawled)"]
        if is_paywall:
            head.append("paywall_warning: Content may be behind a paywall (partial content returned)")
    return "\n".join(head) + "\n\n" + text
