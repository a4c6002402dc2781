"""High-level client: search / crawl / fetch / suggest / stats without MCP (reference infomesh/sdk/client.py:27-279).

``gpu=True`` attaches the HBM-resident index so ``search`` and ``search_many`` run through the fused device pipeline;
otherwise the client uses the SQLite FTS plane exactly like the reference.  The crawl path goes through the real crawl
worker (robots, SSRF, dedup) instead of a bare HTTP GET."""
from __future__ import annotations

import asyncio
from dataclasses import dataclass, replace
from pathlib import Path
from typing import Any


@dataclass
class SearchResult:
    title: str
    url: str
    snippet: str
    score: float
    crawled_at: float = 0.0

    def to_dict(self) -> dict[str, object]:
        return {"title": self.title, "url": self.url, "snippet": self.snippet, "score": self.score, "crawled_at": self.crawled_at}


@dataclass
class CrawlResult:
    url: str
    success: bool
    title: str = ""
    word_count: int = 0
    error: str = ""


@dataclass
class NetworkInfo:
    peer_count: int = 0
    index_size: int = 0
    credit_balance: float = 0.0
    uptime_hours: float = 0.0


class InfoMeshClient:
    def __init__(self, data_dir: str = "~/.infomesh", config: dict[str, Any] | None = None, *, gpu: bool = False):
        self._data_dir = Path(data_dir).expanduser()
        self._overrides = config or {}
        self._gpu = gpu
        self._ctx: Any = None
        self._gpu_index: Any = None

    # ------------------------------------------------------------------ lifecycle
    def _ensure_init(self) -> None:
        if self._ctx is not None:
            return
        from infomesh_b200.config import Config, set_config_value
        from infomesh_b200.services import AppContext

        base = Config()
        cfg = replace(base, node=replace(base.node, data_dir=self._data_dir),
                      index=replace(base.index, db_path=self._data_dir / "index.db", vector_search=False))
        for key, value in self._overrides.items():            # {"crawl.politeness_delay": 0.5, ...}
            cfg = set_config_value(cfg, key, str(value))
        self._data_dir.mkdir(parents=True, exist_ok=True)
        self._ctx = AppContext(cfg)
        if self._gpu:
            from infomesh_b200.engine.multigpu import make_index

            from infomesh_b200.engine.multigpu import warm_start

            self._gpu_index = make_index(self._ctx.store, getattr(cfg, "gpu", None))
            warm_start(self._gpu_index, self._ctx.store, getattr(getattr(cfg, "gpu", None), "segments_dir", ""))

    def close(self) -> None:
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None
        if self._gpu_index is not None and hasattr(self._gpu_index, "close"):
            self._gpu_index.close()
        self._gpu_index = None

    def __enter__(self) -> "InfoMeshClient":
        self._ensure_init()
        return self

    def __exit__(self, *exc: object) -> None:
        self.close()

    # ------------------------------------------------------------------ search
    def search(self, query: str, *, limit: int = 10, offset: int = 0, language: str | None = None,
               include_domains: list[str] | None = None, exclude_domains: list[str] | None = None) -> list[SearchResult]:
        self._ensure_init()
        plain = offset == 0 and not (language or include_domains or exclude_domains)
        if self._gpu_index is not None and plain:
            return [SearchResult(str(h["title"]), str(h["url"]), str(h["snippet"]), float(h["score"]), float(h.get("crawled_at", 0.0)))
                    for h in self._gpu_index.search(query, limit)]
        from infomesh_b200.search.query import search_local

        lg = self._ctx.link_graph
        res = search_local(self._ctx.store, query, limit=limit, offset=offset, language=language, include_domains=include_domains,
                           exclude_domains=exclude_domains, authority_fn=lg.url_authority if lg else None)
        return [SearchResult(r.title, r.url, r.snippet, r.combined_score, r.crawled_at) for r in res.results]

    def search_many(self, queries: list[str], *, limit: int = 10) -> list[list[SearchResult]]:
        """Batch search: one device pass on GPU clients, a loop otherwise."""
        self._ensure_init()
        if self._gpu_index is not None:
            return [[SearchResult(str(h["title"]), str(h["url"]), str(h["snippet"]), float(h["score"]), float(h.get("crawled_at", 0.0))) for h in hits]
                    for hits in self._gpu_index.search_many(queries, limit)]
        return [self.search(q, limit=limit) for q in queries]

    async def search_async(self, query: str, **kwargs: Any) -> list[SearchResult]:
        return await asyncio.to_thread(self.search, query, **kwargs)

    # ------------------------------------------------------------------ crawl / fetch
    async def crawl_async(self, url: str, *, depth: int = 0, force: bool = False) -> CrawlResult:
        from infomesh_b200.services import crawl_and_index

        self._ensure_init()
        if self._ctx.worker is None:
            return CrawlResult(url, False, error="crawler_unavailable")
        try:
            ci = await crawl_and_index(url, worker=self._ctx.worker, store=self._ctx.store, vector_store=self._ctx.vector_store,
                                       link_graph=self._ctx.link_graph, depth=depth, force=force)
        except Exception as exc:  # noqa: BLE001
            return CrawlResult(url, False, error=str(exc))
        if ci.success and self._gpu_index is not None:
            self._gpu_index.note_added()
        doc = self._ctx.store.get_document_by_url(url) if ci.success else None
        return CrawlResult(url, ci.success, ci.title, len(doc.text.split()) if doc else 0, ci.error or "")

    def crawl(self, url: str, *, depth: int = 0, force: bool = False) -> CrawlResult:
        return asyncio.run(self.crawl_async(url, depth=depth, force=force))

    def fetch_page(self, url: str) -> str:
        self._ensure_init()
        doc = self._ctx.store.get_document_by_url(url)
        if doc is None and self.crawl(url).success:
            doc = self._ctx.store.get_document_by_url(url)
        return doc.text if doc is not None else ""

    def add_document(self, url: str, title: str, text: str, *, language: str | None = None) -> int | None:
        """Index text you already have (no crawl)."""
        from infomesh_b200.crawler.parser import ParsedPage
        from infomesh_b200.hashing import content_hash
        from infomesh_b200.services import index_document

        self._ensure_init()
        page = ParsedPage(url=url, title=title, text=text, language=language, raw_html_hash=content_hash(url + text[:64]), text_hash=content_hash(text))
        doc_id = index_document(page, self._ctx.store, self._ctx.vector_store)
        if doc_id is not None and self._gpu_index is not None:
            self._gpu_index.note_added()
        return doc_id

    def refresh_gpu_index(self) -> int:
        """Make newly added documents searchable on the device: incremental append where the index supports it
        (``GpuSearchIndex.refresh``), a rebuild otherwise.  Returns the number of documents appended / indexed."""
        self._ensure_init()
        if self._gpu_index is None:
            return 0
        return getattr(self._gpu_index, "refresh", self._gpu_index.rebuild)()

    # ------------------------------------------------------------------ misc
    def suggest(self, prefix: str, *, limit: int = 5) -> list[str]:
        self._ensure_init()
        return list(self._ctx.store.suggest(prefix, limit=limit))

    def get_stats(self) -> dict[str, object]:
        self._ensure_init()
        out: dict[str, object] = dict(self._ctx.store.get_stats())
        out["total_documents"] = out.get("document_count", 0)
        if self._gpu_index is not None:
            out["gpu"] = self._gpu_index.stats()
        return out

    def network_info(self) -> NetworkInfo:
        import json
        import time

        self._ensure_init()
        info = NetworkInfo(index_size=int(self._ctx.store.get_stats().get("document_count", 0)))
        if self._ctx.ledger is not None:
            info.credit_balance = float(self._ctx.ledger.balance())
        try:
            st = json.loads((self._data_dir / "p2p_status.json").read_text())
            if time.time() - float(st.get("timestamp", 0)) < 30:
                info.peer_count = int(st.get("peers", 0))
        except (OSError, ValueError):
            pass
        from infomesh_b200.runtime import read_runtime_status

        rt = read_runtime_status(self._data_dir)
        info.uptime_hours = float(rt.get("uptime_seconds", 0.0) or 0.0) / 3600
        return info
