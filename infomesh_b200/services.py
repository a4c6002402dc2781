"""Service layer shared by the CLI, MCP server, HTTP API and dashboard: the one crawl -> index -> vector-index ->
publish path, cached page fetches, and :class:`AppContext`, the factory that wires every component from a
:class:`Config` according to the node role (reference infomesh/services.py:36-801)."""
from __future__ import annotations

import contextlib
import time
from dataclasses import dataclass
from typing import Any

from infomesh_b200.config import Config, NodeRole, load_config
from infomesh_b200.crawler.parser import ParsedPage
from infomesh_b200.index.local_store import LocalStore
from infomesh_b200.security import SSRFError, validate_url
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_PAYWALL_SIGNALS = ("subscribe to continue", "sign in to read", "create a free account", "this content is for subscribers")


def is_paywall_content(text: str) -> bool:
    low = text.lower()
    return any(sig in low for sig in _PAYWALL_SIGNALS)


def _truncate_to_bytes(text: str, max_bytes: int) -> str:
    raw = text.encode("utf-8")
    return text if len(raw) <= max_bytes else raw[:max_bytes].decode("utf-8", errors="ignore")


def index_document(page: ParsedPage, store: LocalStore, vector_store: Any | None = None, *, js_required: bool = False) -> int | None:
    """The single place a crawled page enters the indexes.  Returns the doc id, or None for a duplicate."""
    doc_id = store.add_document(url=page.url, title=page.title, text=page.text, raw_html_hash=page.raw_html_hash,
                                text_hash=page.text_hash, language=page.language, js_required=js_required)
    if vector_store is not None and doc_id is not None:
        vector_store.add_document(doc_id=doc_id, url=page.url, title=page.title, text=page.text, language=page.language)
    return doc_id


class _NetworkPublisher:
    """Adapter over whatever can announce documents to the mesh: a running P2P node (preferred) or a bare
    ``DistributedIndex``.  Resolves the callables once; every call is best-effort and reports keywords published."""

    def __init__(self, p2p_node: Any | None, distributed_index: Any | None):
        self._one = self._many = None
        self._one_takes_kwargs = False
        node_one = getattr(p2p_node, "publish_document_to_network", None)
        if callable(node_one):
            self._one = node_one
        else:
            index_one = getattr(distributed_index, "publish_document", None)
            if callable(index_one):
                self._one, self._one_takes_kwargs = index_one, True
        for owner, name in ((p2p_node, "publish_documents_to_network"), (distributed_index, "publish_batch")):
            fn = getattr(owner, name, None)
            if callable(fn):
                self._many = fn
                break
        self.available = p2p_node is not None or distributed_index is not None

    @staticmethod
    def _count(value: Any) -> int:
        return value if isinstance(value, int) else 0

    async def one(self, page: ParsedPage, doc_id: int) -> int:
        if self._one is None:
            return 0
        if self._one_takes_kwargs:
            return self._count(await self._one(doc_id=doc_id, url=page.url, title=page.title, text=page.text))
        return self._count(await self._one(doc_id, page.url, page.title, page.text))

    async def many(self, docs: list[Any]) -> int:
        return self._count(await self._many(docs)) if self._many is not None else 0


async def publish_document_to_network(page: ParsedPage, doc_id: int | None, *, p2p_node: Any | None = None,
                                      distributed_index: Any | None = None) -> int:
    """Announce one freshly indexed page; duplicates (``doc_id is None``) are never announced.  Never raises."""
    if doc_id is None:
        return 0
    try:
        return await _NetworkPublisher(p2p_node, distributed_index).one(page, doc_id)
    except Exception as exc:  # noqa: BLE001
        logger.warning("distributed_publish_failed", url=page.url, error=str(exc))
        return 0


def _publish_windows(store: LocalStore, batch_size: int, limit: int | None):
    """Yield ``(offset_after, docs)`` windows over the store until it is exhausted or ``limit`` documents were read."""
    seen = 0
    while limit is None or seen < limit:
        want = batch_size if limit is None else min(batch_size, limit - seen)
        docs = store.get_documents_for_publish(limit=want, offset=seen)
        if not docs:
            return
        seen += len(docs)
        yield seen, docs


async def republish_local_index(store: LocalStore, *, p2p_node: Any | None = None, distributed_index: Any | None = None,
                                batch_size: int = 250, limit: int | None = None) -> int:
    """Walk the local index in windows of ``batch_size`` (clamped to 1..1000) and re-announce every document, e.g. after
    a restart.  A failing window is logged and skipped.  Returns the keywords published."""
    publisher = _NetworkPublisher(p2p_node, distributed_index)
    if not publisher.available:
        return 0
    scanned = published = 0
    for scanned, docs in _publish_windows(store, max(1, min(batch_size, 1000)), limit):
        try:
            published += await publisher.many(docs)
        except Exception as exc:  # noqa: BLE001
            logger.warning("distributed_republish_batch_failed", offset=scanned, error=str(exc))
    logger.info("distributed_index_republished", documents_scanned=scanned, keywords_published=published)
    return published


@dataclass(frozen=True)
class FetchPageResult:
    success: bool
    title: str = ""
    url: str = ""
    text: str = ""
    is_cached: bool = False
    is_stale: bool = False
    is_paywall: bool = False
    crawled_at: float = 0.0
    error: str | None = None


class _PageFetcher:
    """cache -> (optionally) crawl -> classify, with the size cap and staleness rule in one place."""

    _PAYWALL_STATUS = frozenset({"http_402", "http_403"})

    def __init__(self, store: LocalStore, vector_store: Any | None, max_size_bytes: int, cache_ttl_seconds: int):
        self.store, self.vector_store = store, vector_store
        self.cap, self.ttl = max_size_bytes, cache_ttl_seconds

    def failure(self, url: str, error: str | None, **flags: Any) -> FetchPageResult:
        return FetchPageResult(False, url=url, error=error, **flags)

    def from_cache(self, url: str) -> FetchPageResult:
        try:
            validate_url(url)
        except SSRFError as exc:
            return self.failure(url, f"blocked: {exc}")
        doc = self.store.get_document_by_url(url)
        if doc is None:
            return self.failure(url, "not_cached")
        age = time.time() - doc.crawled_at
        return FetchPageResult(True, title=doc.title, url=doc.url, text=_truncate_to_bytes(doc.text, self.cap), is_cached=True,
                               is_stale=age > self.ttl, crawled_at=doc.crawled_at)

    async def from_network(self, url: str, worker: Any) -> FetchPageResult:
        if worker is None:
            return self.failure(url, "crawler_unavailable")
        outcome = await worker.crawl_url(url)
        page = outcome.page if outcome.success else None
        if page:
            index_document(page, self.store, self.vector_store, js_required=outcome.js_required)
            return FetchPageResult(True, title=page.title, url=url, text=_truncate_to_bytes(page.text, self.cap),
                                   is_paywall=is_paywall_content(page.text), crawled_at=time.time())
        if outcome.error in self._PAYWALL_STATUS:
            return self.failure(url, f"paywall:{outcome.error}", is_paywall=True)
        return self.failure(url, outcome.error)


def fetch_page(url: str, *, store: LocalStore, worker: Any = None, vector_store: Any | None = None,
               max_size_bytes: int = 102_400, cache_ttl_seconds: int = 604_800) -> FetchPageResult:
    """Cache-only lookup (sync).  ``error="not_cached"`` tells the caller to crawl."""
    return _PageFetcher(store, vector_store, max_size_bytes, cache_ttl_seconds).from_cache(url)


async def fetch_page_async(url: str, *, store: LocalStore, worker: Any, vector_store: Any | None = None,
                           max_size_bytes: int = 102_400, cache_ttl_seconds: int = 604_800) -> FetchPageResult:
    """Cached copy if there is one (even a stale one, flagged), else crawl + index + return; SSRF refusals are final."""
    fetcher = _PageFetcher(store, vector_store, max_size_bytes, cache_ttl_seconds)
    cached = fetcher.from_cache(url)
    if cached.success or (cached.error or "").startswith("blocked"):
        return cached
    return await fetcher.from_network(url, worker)


@dataclass(frozen=True)
class CrawlAndIndexResult:
    success: bool
    title: str = ""
    url: str = ""
    text_length: int = 0
    links_discovered: int = 0
    elapsed_ms: float = 0.0
    error: str | None = None


async def crawl_and_index(url: str, *, worker: Any, store: LocalStore, vector_store: Any | None = None,
                          p2p_node: Any | None = None, distributed_index: Any | None = None, link_graph: Any | None = None,
                          depth: int = 0, force: bool = False) -> CrawlAndIndexResult:
    res = await worker.crawl_url(url, depth=depth, force=force)
    if not (res.success and res.page):
        return CrawlAndIndexResult(False, url=url, error=res.error)
    if link_graph is not None and res.discovered_links:
        try:
            link_graph.add_links(url, res.discovered_links)
        except Exception as exc:  # noqa: BLE001
            logger.warning("link_graph_update_failed", url=url, error=str(exc))
    try:
        doc_id = index_document(res.page, store, vector_store, js_required=res.js_required)
        await publish_document_to_network(res.page, doc_id, p2p_node=p2p_node, distributed_index=distributed_index)
    except Exception as exc:  # noqa: BLE001
        logger.error("index_document_failed", url=url, error=str(exc))
        return CrawlAndIndexResult(False, url=url, error="index_failed")
    return CrawlAndIndexResult(True, res.page.title, url, len(res.page.text), len(res.discovered_links), res.elapsed_ms)


class AppContext:
    """Everything a process needs, built from config and gated on ``node.role``:

    * full / crawler -> dedup DB, robots cache, scheduler, crawl worker (+ feed monitor & priority recrawl queue)
    * full / search  -> link graph, credit ledger, vector store (GPU-resident when a device is present), feedback
    * crawler with submit peers -> IndexSubmitSender;  search -> IndexSubmitReceiver
    * ``[llm] enabled`` -> summariser backend;  GitHub identity -> cross-node credit sync
    """

    def __init__(self, config: Config | None = None, *, apply_os_priority: bool = False):
        from infomesh_b200.credits.github_identity import resolve_github_email
        from infomesh_b200.p2p.keys import ensure_keys
        from infomesh_b200.resources.governor import ResourceGovernor
        from infomesh_b200.resources.profiles import get_profile

        self.config = c = config or load_config()
        role = str(c.node.role).lower()          # the loader keeps the value as written ("FULL" is valid)
        crawls, searches = role in (NodeRole.FULL, NodeRole.CRAWLER), role in (NodeRole.FULL, NodeRole.SEARCH)
        self.governor = ResourceGovernor(get_profile(c.resources.profile))
        if apply_os_priority:
            self.governor.apply_os_priority()
        self.github_email: str = resolve_github_email(c) or ""
        self.store = LocalStore(db_path=c.index.db_path, tokenizer=c.index.fts_tokenizer,
                                compression_enabled=c.storage.compression_enabled, compression_level=c.storage.compression_level)
        self.key_pair = None
        try:
            self.key_pair = ensure_keys(c.node.data_dir)
        except Exception as exc:  # noqa: BLE001
            logger.warning("key_init_failed", error=str(exc))

        self.dedup = self.robots = self.scheduler = self.worker = None
        self.feed_monitor = self.priority_queue = None
        if crawls:
            from infomesh_b200.crawler.dedup import DeduplicatorDB
            from infomesh_b200.crawler.robots import RobotsChecker
            from infomesh_b200.crawler.scheduler import Scheduler
            from infomesh_b200.crawler.worker import CrawlWorker

            self.dedup = DeduplicatorDB(str(c.node.data_dir / "dedup.db"))
            self.robots = RobotsChecker(c.crawl.user_agent)
            self.scheduler = Scheduler(politeness_delay=c.crawl.politeness_delay, urls_per_hour=c.crawl.urls_per_hour,
                                       pending_per_domain=c.crawl.pending_per_domain, max_depth=c.crawl.max_depth)
            self.worker = CrawlWorker(c.crawl, self.scheduler, self.dedup, self.robots)
            if getattr(c.crawl, "rss_enabled", False):
                from infomesh_b200.crawler.feed_monitor import FeedMonitor
                from infomesh_b200.crawler.freshness import PriorityRecrawlQueue

                self.feed_monitor, self.priority_queue = FeedMonitor(), PriorityRecrawlQueue()

        self.link_graph = self.ledger = self.vector_store = self.feedback_store = None
        if searches:
            from infomesh_b200.index.link_graph import LinkGraph

            self.link_graph = LinkGraph(str(c.node.data_dir / "links.db"))
            try:
                from infomesh_b200.credits.ledger import CreditLedger

                self.ledger = CreditLedger(c.node.data_dir / "credits.db", owner_email=self.github_email)
            except Exception as exc:  # noqa: BLE001
                logger.warning("ledger_init_failed", error=str(exc))
            if c.index.vector_search:
                try:
                    from infomesh_b200.index.vector_store import VectorStore

                    self.vector_store = VectorStore(persist_dir=c.node.data_dir / "vectors", model_name=c.index.embedding_model)
                except Exception as exc:  # noqa: BLE001
                    logger.warning("vector_search_unavailable", reason=str(exc))
            try:
                from infomesh_b200.search.feedback import FeedbackStore

                self.feedback_store = FeedbackStore(str(c.node.data_dir / "feedback.db"))
            except Exception as exc:  # noqa: BLE001
                logger.debug("feedback_store_unavailable", error=str(exc))

        self.llm_backend = None
        if c.llm.enabled:
            try:
                from infomesh_b200.summarizer.engine import create_backend

                self.llm_backend = create_backend(c.llm.runtime, c.llm.model)
            except Exception as exc:  # noqa: BLE001
                logger.warning("llm_backend_unavailable", error=str(exc))

        self.index_submit_sender = self.index_submit_receiver = None
        if role == NodeRole.CRAWLER and c.network.index_submit_peers:
            from infomesh_b200.p2p.index_submit import IndexSubmitSender

            self.index_submit_sender = IndexSubmitSender(c, self.key_pair)
        if role == NodeRole.SEARCH:
            from infomesh_b200.p2p.index_submit import IndexSubmitReceiver

            self.index_submit_receiver = IndexSubmitReceiver(c, self.store, self.vector_store, self.key_pair)

        self.credit_sync_manager = None
        if self.ledger is not None and self.github_email:
            try:
                from infomesh_b200.credits.sync import CreditSyncManager, CreditSyncStore

                self.credit_sync_manager = CreditSyncManager(self.ledger, CreditSyncStore(c.node.data_dir / "credit_sync.db"),
                                                             self.github_email, key_pair=self.key_pair,
                                                             local_peer_id=self.key_pair.peer_id if self.key_pair else "")
            except Exception as exc:  # noqa: BLE001
                logger.warning("credit_sync_init_failed", error=str(exc))
        self.governor.check_and_adjust()
        self.distributed_index = None
        self.p2p_node = None
        logger.info("app_context_initialized", role=str(role), has_crawler=self.worker is not None,
                    has_search=self.link_graph is not None, has_vector=self.vector_store is not None,
                    has_credit_sync=self.credit_sync_manager is not None)

    # ------------------------------------------------------------------ teardown
    def close(self) -> None:
        if self.p2p_node is not None:
            with contextlib.suppress(Exception):
                self.p2p_node.stop()
        for obj in (self.credit_sync_manager, self.feedback_store, self.vector_store, self.ledger, self.link_graph, self.store,
                    self.dedup):
            if obj is not None:
                with contextlib.suppress(Exception):
                    obj.close()

    async def close_async(self) -> None:
        if self.worker is not None:
            with contextlib.suppress(Exception):
                await self.worker.close()
        if self.llm_backend is not None:
            with contextlib.suppress(Exception):
                await self.llm_backend.close()
        self.close()

    def __enter__(self) -> "AppContext":
        return self

    def __exit__(self, *exc: object) -> None:
        self.close()

    async def __aenter__(self) -> "AppContext":
        return self

    async def __aexit__(self, *exc: object) -> None:
        await self.close_async()


def create_local_search_fn(config: Config, store: LocalStore | None = None):
    """Async ``(query, limit) -> [result dicts]`` used to answer inbound peer searches.  Opens its own read
    connection (WAL allows concurrent readers) unless a store is passed."""
    try:
        import asyncio

        from infomesh_b200.search.query import search_local

        ls = store or LocalStore(db_path=config.index.db_path, tokenizer=config.index.fts_tokenizer,
                                 compression_enabled=config.storage.compression_enabled,
                                 compression_level=config.storage.compression_level)

        async def _local_search(query: str, limit: int = 10) -> list[dict[str, object]]:
            qr = await asyncio.to_thread(search_local, ls, query, limit=limit)
            return [{"url": r.url, "title": r.title, "snippet": r.snippet, "score": r.combined_score, "doc_id": r.doc_id}
                    for r in qr.results]

        return _local_search
    except Exception as exc:  # noqa: BLE001
        logger.warning("local_search_fn_unavailable", error=str(exc))
        return None


def bootstrap_p2p(config: Config, *, credit_sync_manager: Any | None = None, local_search_fn: Any | None = None,
                  store_fn: Any | None = None, index_submit_receiver: Any | None = None, llm_handler: Any | None = None,
                  **node_kwargs: Any) -> tuple[Any | None, Any | None]:
    """Best effort: (node, distributed_index), or (None, None) with the process continuing in local-only mode."""
    try:
        from infomesh_b200.p2p.node import InfoMeshNode

        node = InfoMeshNode(config, credit_sync_manager=credit_sync_manager, local_search_fn=local_search_fn,
                            store_fn=store_fn, index_submit_receiver=index_submit_receiver, llm_handler=llm_handler,
                            **node_kwargs)
        node.start(blocking=False)
        logger.info("p2p_started", peer_id=node.peer_id, listen_port=config.node.listen_port,
                    bootstrap_nodes=len(config.network.bootstrap_nodes))
        if not config.network.bootstrap_nodes:
            logger.warning("p2p_no_bootstrap", msg="No bootstrap nodes configured: add [network] bootstrap_nodes to config.toml")
    except Exception as exc:  # noqa: BLE001
        logger.error("p2p_start_failed", error=str(exc), msg="running in local-only mode")
        return None, None
    return node, node.distributed_index


def __getattr__(name: str):
    # the crawl loop lives in crawler.crawl_loop; re-exported lazily to avoid an import cycle
    if name in ("seed_and_crawl_loop", "_reseed_queue"):
        from infomesh_b200.crawler import crawl_loop

        return getattr(crawl_loop, name)
    raise AttributeError(name)
