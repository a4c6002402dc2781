"""search/query.py (local / hybrid / distributed), merge.py (RRF), cache.py, formatter.py over an in-memory LocalStore."""
import asyncio
import json
import time
from types import SimpleNamespace

import pytest

from infomesh_b200.hashing import content_hash
from infomesh_b200.index.local_store import LocalStore
from infomesh_b200.search import formatter as FMT
from infomesh_b200.search import query as Q
from infomesh_b200.search.cache import QueryCache
from infomesh_b200.search.merge import merge_results

DOCS = [
    ("https://docs.python.org/asyncio", "Python asyncio guide", "The asyncio event loop runs coroutines and tasks. " * 8, "en"),
    ("https://ex.org/rust-tokio", "Rust tokio runtime", "Tokio is an asynchronous runtime for Rust with a scheduler. " * 8, "en"),
    ("https://blog.io/cooking", "Cooking pasta", "Boil water, add salt, cook the pasta until al dente. " * 8, "en"),
    ("https://ex.kr/search", "검색 엔진 소개", "분산 검색 엔진은 여러 노드가 색인을 나누어 가진다. " * 8, "ko"),
    ("https://ex.org/db", "Database errors", "A datastore exception can corrupt the storage layer in rare cases. " * 8, "en"),
]


@pytest.fixture()
def store():
    s = LocalStore(None)
    for url, title, text, lang in DOCS:
        s.add_document(url, title, text, content_hash("raw" + url), content_hash(text), language=lang)
    yield s
    s.close()


# ------------------------------------------------------------------ sanitising
@pytest.mark.parametrize("raw,clean", [
    ('python "asyncio" (loop)*', "python asyncio loop"), ("a AND b OR NOT c NEAR d", "a b c d"), ('"""***', "infomesh"),
    ("col:value ^boost {x}", "col value boost x"), ("   spaced    out   ", "spaced out")])
def test_sanitize_fts_query(raw, clean):
    assert Q.sanitize_fts_query(raw) == clean


def test_sanitize_caps_length():
    assert len(Q.sanitize_fts_query("word " * 1000)) <= 1000


# ------------------------------------------------------------------ local search
def test_search_local_ranks_and_filters(store):
    r = Q.search_local(store, "asyncio event loop")
    assert r.source == "local" and r.results[0].url == DOCS[0][0] and r.total == len(r.results) and r.elapsed_ms >= 0
    assert "<b>" in r.results[0].snippet or "asyncio" in r.results[0].snippet.lower()
    assert Q.search_local(store, "asyncio", language="ko").total == 0
    assert Q.search_local(store, "asyncio", exclude_domains=["docs.python.org"]).total == 0
    assert Q.search_local(store, "asyncio", include_domains=["docs.python.org"]).total == 1
    assert Q.search_local(store, "asyncio", date_from=time.time() + 100).total == 0
    assert Q.search_local(store, "zzzznotaword").total == 0


def test_search_local_widens_sparse_queries_with_synonyms(store):
    r = Q.search_local(store, "database error")                    # literal AND matches the title; synonyms find the body too
    assert any(x.url == "https://ex.org/db" for x in r.results)
    r2 = Q.search_local(store, "db")                                # only reachable through expand_query("db") -> database / datastore
    assert [x.url for x in r2.results] == ["https://ex.org/db"]


def test_search_local_cjk_query_uses_bigrams(store):
    s = LocalStore(None, tokenizer="trigram")
    s.add_document(DOCS[3][0], DOCS[3][1], DOCS[3][2], "r", content_hash(DOCS[3][2]), language="ko")
    assert Q.search_local(s, "검색 엔진").total >= 0                 # must not raise on FTS syntax
    s.close()


def test_authority_function_changes_order(store):
    base = Q.search_local(store, "runtime OR asyncio scheduler tasks")
    boosted = Q.search_local(store, "asynchronous runtime", authority_fn=lambda u: 1.0 if "tokio" in u else 0.0)
    assert boosted.results and boosted.results[0].url == DOCS[1][0] and boosted.results[0].authority_score == 1.0
    assert isinstance(base.results, list)


# ------------------------------------------------------------------ RRF merge / hybrid
def hit(url, score, doc_id=1, title="T", snippet="snip", preview="prev"):
    return SimpleNamespace(url=url, score=score, doc_id=doc_id, title=title, snippet=snippet, text_preview=preview)


def test_rrf_merge_scores_and_sources():
    fts = [hit("a", 9.0), hit("b", 5.0), hit("a", 1.0)]
    vec = [hit("b", 0.9, title="", snippet=""), hit("c", 0.8)]
    m = merge_results(fts, vec, rrf_k=60)
    by = {x.url: x for x in m}
    assert by["b"].source == "hybrid" and by["b"].combined_score == round(1 / 62 + 1 / 61, 6) and m[0].url == "b"
    assert by["a"].source == "fts" and by["a"].fts_score == 9.0 and by["a"].combined_score == round(1 / 61, 6)
    assert by["c"].source == "vector" and by["c"].snippet == "prev" and by["c"].vector_score == 0.8
    heavy = merge_results(fts, vec, vector_weight=10.0)
    assert [x.url for x in heavy][:2] == ["b", "c"] and len(merge_results(fts, vec, limit=1)) == 1


class FakeVectors:
    def __init__(self, hits):
        self.hits = hits

    def search(self, query, limit=10):
        return self.hits[:limit]

    def add_document(self, *a, **k):
        pass


def test_search_hybrid_sources_and_type_check(store):
    vec = FakeVectors([hit(DOCS[0][0], 0.9, doc_id=1), hit("https://semantic.only/x", 0.7, doc_id=99)])
    h = Q.search_hybrid(store, vec, "asyncio")
    assert h.source == "hybrid" and h.results[0].url == DOCS[0][0] and h.results[0].source == "hybrid"
    assert any(r.source == "vector" for r in h.results)
    assert Q.search_hybrid(store, FakeVectors([]), "asyncio").source == "fts"
    assert Q.search_hybrid(store, vec, "zzzznotaword").source == "vector"
    with pytest.raises(TypeError):
        Q.search_hybrid(store, object(), "asyncio")


# ------------------------------------------------------------------ distributed
def test_search_distributed_merges_remote_and_survives_failures(store):
    async def net(query, keywords, limit):
        return [{"url": "https://peer.org/p", "title": "Peer", "snippet": "s", "score": "7.5", "doc_id": "12", "peer_id": "p1"},
                {"url": DOCS[0][0], "title": "dup", "score": 99.0, "peer_id": "p1"}, {"title": "no url"}, "junk",
                {"url": "https://peer.org/nan", "score": float("nan"), "doc_id": True}]

    r = asyncio.run(Q.search_distributed(store, None, "asyncio event", network_search_fn=net))
    assert r.source == "distributed" and r.remote_count == 3 and r.local_count >= 1
    urls = [x.url for x in r.results]
    assert urls[0] == "https://peer.org/p" and urls.count(DOCS[0][0]) == 1                      # local copy wins the duplicate
    remote = r.results[0]
    assert remote.combined_score == 7.5 and remote.doc_id == 12 and remote.peer_id == "p1" and remote.bm25_score == 0.0
    assert next(x for x in r.results if x.url.endswith("/nan")).combined_score == 0.0

    async def boom(*a):
        raise OSError("network down")

    r2 = asyncio.run(Q.search_distributed(store, None, "asyncio", network_search_fn=boom))
    assert r2.source == "local_only" and r2.remote_count == 0 and r2.results

    class Ptrs:
        async def query(self, kws):
            return [SimpleNamespace(url="https://dht.org/x", title="D", score=0.3, doc_id=5, peer_id="p9")]

    r3 = asyncio.run(Q.search_distributed(store, Ptrs(), "asyncio"))
    assert r3.source == "distributed" and any(x.peer_id == "p9" for x in r3.results)


# ------------------------------------------------------------------ cache
def test_query_cache_lru_ttl_and_keys():
    c = QueryCache(max_size=2, ttl_seconds=10)
    k1, k2, k3 = (QueryCache.make_key(q) for q in ("A ", "b", "c"))
    assert k1 == QueryCache.make_key("a") and k1 != QueryCache.make_key("a", 20) and QueryCache.make_key("a", lang="ko") != k1
    c.put(k1, 1, now=0), c.put(k2, 2, now=1)
    assert c.get(k1, now=2) == 1                                   # refreshes recency
    c.put(k3, 3, now=3)                                            # evicts k2
    assert c.get(k2, now=3) is None and c.get(k3, now=3) == 3 and len(c) == 2
    assert c.get(k1, now=11) is None                               # TTL from insertion, not from access
    st = c.stats
    assert (st.hits, st.misses, st.evictions) == (2, 2, 2) and st.hit_rate == 0.5 and st.total == 4     # expiry counts as an eviction
    assert c.invalidate(k3) and not c.invalidate(k3)
    c.invalidate()
    assert len(c) == 0


def test_query_cache_reference_calling_convention():
    c = QueryCache(max_size=3, ttl_seconds=10)
    c.put("Rust Ownership ", 10, ["r1", "r2"], now=0)
    assert c.get("rust ownership", 10, now=1) == ["r1", "r2"] and c.get("rust ownership", 20, now=1) is None
    assert c.get(QueryCache.make_key("rust ownership", 10), now=1) == ["r1", "r2"]          # both conventions share one store
    c.put("old", 5, ["x"], now=0), c.put("new", 5, ["y"], now=9)
    assert c.size == 3 and c.evict_expired(now=10.5) == 2 and c.size == 1 and c.get("new", 5, now=10.5) == ["y"]
    assert c.invalidate("new", 5) is True and c.invalidate("new", 5) is False
    c.put("a", 1, []), c.clear()
    assert c.size == 0 and c.stats.evictions == 2


# ------------------------------------------------------------------ formatting
def test_formatters_text_and_json(store):
    r = Q.search_local(store, "asyncio")
    text = FMT.format_fts_results(r)
    assert DOCS[0][0] in text and "Python asyncio guide" in text
    data = json.loads(FMT.format_fts_results_json(r, max_snippet=20))
    assert data["total"] == r.total and data["results"][0]["url"] == DOCS[0][0] and len(data["results"][0]["snippet"]) <= 23
    h = Q.search_hybrid(store, FakeVectors([hit(DOCS[0][0], 0.9)]), "asyncio")
    assert json.loads(FMT.format_hybrid_results_json(h))["source"] == "hybrid" and "hybrid" in FMT.format_hybrid_results(h).lower()
    empty = Q.search_local(store, "zzzznotaword")
    assert "No results" in FMT.format_fts_results(empty) or FMT.format_fts_results(empty)
    page = FMT.format_fetch_result(title="T", url="https://u", text="body " * 1000, is_cached=True, crawled_at=time.time() - 7200)
    assert "https://u" in page and "T" in page
