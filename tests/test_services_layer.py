"""Service layer: the single index path, cached fetches, crawl-and-index, network publishing and role-gated wiring
(model: reference tests/test_services.py)."""
import asyncio
import time
from dataclasses import replace

import pytest

from infomesh_b200 import services as S
from infomesh_b200.config import Config, NodeRole
from infomesh_b200.crawler.parser import ParsedPage
from infomesh_b200.crawler.worker import CrawlResult
from infomesh_b200.index.local_store import LocalStore


def _page(url="https://example.org/a", text="Programmatic dependent launch lets the next kernel's prologue overlap this kernel's tail. " * 4, h="1"):
    return ParsedPage(url=url, title="PDL", text=text, language="en", raw_html_hash="r" + h, text_hash="t" + h)


class FakeWorker:
    def __init__(self, result):
        self.result, self.calls = result, []

    async def crawl_url(self, url, depth=0, force=False):
        self.calls.append((url, depth, force))
        return self.result


class FakeVectors:
    def __init__(self):
        self.added = []

    def add_document(self, **kw):
        self.added.append(kw)


@pytest.fixture()
def store(tmp_path):
    with LocalStore(tmp_path / "index.db") as st:
        yield st


def test_index_document_feeds_both_indexes_once(store):
    vec = FakeVectors()
    doc_id = S.index_document(_page(), store, vec)
    assert doc_id and vec.added[0]["doc_id"] == doc_id and vec.added[0]["url"] == "https://example.org/a"
    assert S.index_document(_page(), store, vec) is None and len(vec.added) == 1          # duplicate: neither index touched


def test_helpers_detect_paywalls_and_cut_on_byte_boundaries():
    assert S.is_paywall_content("Please Subscribe to continue reading") and not S.is_paywall_content("free article")
    cut = S._truncate_to_bytes("한국어 텍스트", 7)
    assert cut == "한국" and S._truncate_to_bytes("short", 100) == "short"


def test_fetch_page_is_cache_only_and_flags_stale_copies(store):
    S.index_document(_page(), store)
    hit = S.fetch_page("https://example.org/a", store=store, max_size_bytes=64)
    assert hit.success and hit.is_cached and not hit.is_stale and len(hit.text.encode()) <= 64
    assert S.fetch_page("https://example.org/a", store=store, cache_ttl_seconds=-1).is_stale
    assert S.fetch_page("https://example.org/missing", store=store).error == "not_cached"
    assert S.fetch_page("http://169.254.169.254/latest", store=store).error.startswith("blocked")


def test_fetch_page_async_crawls_on_a_miss_and_maps_paywall_codes(store):
    ok = FakeWorker(CrawlResult("https://example.org/new", True, page=_page("https://example.org/new", "Sign in to read the rest of this story. " * 8, "2")))
    got = asyncio.run(S.fetch_page_async("https://example.org/new", store=store, worker=ok))
    assert got.success and not got.is_cached and got.is_paywall and store.get_document_by_url("https://example.org/new") is not None
    again = asyncio.run(S.fetch_page_async("https://example.org/new", store=store, worker=ok))
    assert again.is_cached and len(ok.calls) == 1
    denied = asyncio.run(S.fetch_page_async("https://example.org/pay", store=store, worker=FakeWorker(CrawlResult("x", False, error="http_402"))))
    assert denied.error == "paywall:http_402" and denied.is_paywall
    assert asyncio.run(S.fetch_page_async("https://example.org/x", store=store, worker=FakeWorker(CrawlResult("x", False, error="timeout")))).error == "timeout"
    assert asyncio.run(S.fetch_page_async("https://example.org/x", store=store, worker=None)).error == "crawler_unavailable"
    blocked = FakeWorker(None)
    assert asyncio.run(S.fetch_page_async("http://10.0.0.1/", store=store, worker=blocked)).error.startswith("blocked") and not blocked.calls


def test_crawl_and_index_updates_link_graph_and_publishes(store):
    class Graph:
        def __init__(self):
            self.edges = []

        def add_links(self, src, dst):
            self.edges.append((src, tuple(dst)))

    class Dist:
        def __init__(self):
            self.docs = []

        async def publish_document(self, **kw):
            self.docs.append(kw)
            return 7

    res = CrawlResult("https://example.org/a", True, page=_page(), discovered_links=["https://example.org/b", "https://example.org/c"], elapsed_ms=12.0)
    g, d, w = Graph(), Dist(), FakeWorker(res)
    out = asyncio.run(S.crawl_and_index("https://example.org/a", worker=w, store=store, link_graph=g, distributed_index=d, depth=2, force=True))
    assert out.success and out.links_discovered == 2 and out.text_length == len(_page().text) and out.elapsed_ms == 12.0
    assert w.calls == [("https://example.org/a", 2, True)] and g.edges[0][1] == ("https://example.org/b", "https://example.org/c") and d.docs[0]["doc_id"]
    failed = asyncio.run(S.crawl_and_index("https://example.org/z", worker=FakeWorker(CrawlResult("z", False, error="robots_disallowed")), store=store))
    assert not failed.success and failed.error == "robots_disallowed"

    class BrokenStore:
        def add_document(self, **kw):
            raise RuntimeError("disk full")

    assert asyncio.run(S.crawl_and_index("https://example.org/a", worker=w, store=BrokenStore())).error == "index_failed"


def test_publish_prefers_the_node_and_swallows_transport_errors():
    class Node:
        async def publish_document_to_network(self, doc_id, url, title, text):
            return 11

    class Flaky:
        async def publish_document(self, **kw):
            raise ConnectionError("no peers")

    run = asyncio.run
    assert run(S.publish_document_to_network(_page(), None, p2p_node=Node())) == 0             # duplicates are not published
    assert run(S.publish_document_to_network(_page(), 3, p2p_node=Node(), distributed_index=Flaky())) == 11
    assert run(S.publish_document_to_network(_page(), 3, distributed_index=Flaky())) == 0
    assert run(S.publish_document_to_network(_page(), 3)) == 0


def test_republish_walks_the_store_in_batches(store):
    for i in range(7):
        S.index_document(_page(f"https://example.org/{i}", f"Document number {i} about persistent kernels and tile schedulers. " * 4, str(i)), store)

    class Dist:
        def __init__(self):
            self.sizes = []

        async def publish_batch(self, docs):
            self.sizes.append(len(docs))
            if len(self.sizes) == 2:
                raise TimeoutError("slow peer")
            return len(docs) * 2

    d = Dist()
    assert asyncio.run(S.republish_local_index(store, distributed_index=d, batch_size=3)) == (3 + 1) * 2 and d.sizes == [3, 3, 1]
    d2 = Dist()
    assert asyncio.run(S.republish_local_index(store, distributed_index=d2, batch_size=5, limit=4)) == 8 and d2.sizes == [4]
    assert asyncio.run(S.republish_local_index(store)) == 0


def _cfg(tmp_path, role, **net):
    base = Config()
    return replace(base, node=replace(base.node, data_dir=tmp_path, role=role), index=replace(base.index, db_path=tmp_path / "index.db", vector_search=False),
                   network=replace(base.network, **net))


def test_app_context_wires_components_by_role(tmp_path):
    with S.AppContext(_cfg(tmp_path / "full", NodeRole.FULL)) as full:
        assert full.worker is not None and full.scheduler is not None and full.link_graph is not None and full.ledger is not None
        assert full.index_submit_sender is None and full.index_submit_receiver is None and full.vector_store is None and full.key_pair is not None
    with S.AppContext(_cfg(tmp_path / "crawler", NodeRole.CRAWLER, index_submit_peers=["peer-1"])) as cr:
        assert cr.worker is not None and cr.link_graph is None and cr.ledger is None and cr.feedback_store is None
        assert cr.index_submit_sender is not None and cr.index_submit_sender.submit_peers == ["peer-1"]
    with S.AppContext(_cfg(tmp_path / "crawler2", NodeRole.CRAWLER)) as lone:
        assert lone.index_submit_sender is None
    with S.AppContext(_cfg(tmp_path / "search", NodeRole.SEARCH)) as se:
        assert se.worker is None and se.dedup is None and se.feedback_store is not None and se.index_submit_receiver is not None

    async def go():
        async with S.AppContext(_cfg(tmp_path / "async", NodeRole.FULL)) as ctx:
            return ctx.worker is not None

    assert asyncio.run(go())


def test_local_search_fn_answers_peer_queries(store):
    S.index_document(_page(), store)
    fn = S.create_local_search_fn(Config(), store)
    rows = asyncio.run(fn("dependent launch prologue", limit=3))
    assert rows and set(rows[0]) == {"url", "title", "snippet", "score", "doc_id"} and rows[0]["url"] == "https://example.org/a"


def test_crawl_loop_is_reexported_lazily():
    from infomesh_b200.crawler import crawl_loop

    assert S.seed_and_crawl_loop is crawl_loop.seed_and_crawl_loop
    with pytest.raises(AttributeError):
        S.no_such_symbol
