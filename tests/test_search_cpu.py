"""LocalStore + query pipeline + formatting on CPU (BASELINE config #1 plumbing)."""
from __future__ import annotations

import json
import time

import pytest

from infomesh_b200.hashing import content_hash
from infomesh_b200.index.local_store import LocalStore
from infomesh_b200.index.ranking import (combined_score, freshness_score, normalize_bm25, rank_local_results,
                                         rank_results, RawCandidate)
from infomesh_b200.search import cjk, nlp, passage
from infomesh_b200.search.cache import QueryCache
from infomesh_b200.search.formatter import (format_fetch_result, format_fts_results, format_fts_results_json,
                                            format_hybrid_results)
from infomesh_b200.search.merge import merge_results
from infomesh_b200.search.query import _sanitize_fts_query, search_hybrid, search_local


def test_store_add_search_duplicate_delete(filled_store):
    s = filled_store
    assert s.get_stats()["document_count"] == 4
    assert s.add_document("https://sqlite.org/fts5.html", "dup", "other text", "h", "t2") is None      # url dup
    assert s.add_document("https://new", "dup", "x", "h", content_hash(
        "FTS5 is an SQLite virtual table module that provides full-text search functionality to database "
        "applications. The bm25 ranking function returns a value indicating how well a row matches the query."
    )) is None                                                                                          # text dup
    hits = s.search("asyncio")
    assert hits and hits[0].url.endswith("asyncio.html") and hits[0].score > 0 and "<b>" in hits[0].snippet
    assert s.search("asyncio", include_domains=["sqlite.org"]) == []
    assert s.search("python", exclude_domains=["docs.python.org"])[0].url != "https://docs.python.org/3/library/asyncio.html"
    assert s.search('"unbalanced') == []                       # FTS5 syntax error is swallowed
    doc = s.get_document_by_url("https://www.python-httpx.org/")
    assert doc is not None and s.delete_document(doc.doc_id) and s.search("HTTPX") == []
    assert s.suggest("rust") == ["The Rust Programming Language"]
    assert ("sqlite.org", 1) in s.get_top_domains() and s.get_domain_count() == 3


def test_store_update_recrawl_and_compression(tmp_path):
    s = LocalStore(tmp_path / "i.db", compression_enabled=True)
    did = s.add_document("https://a/b", "T", "alpha beta gamma " * 20, "rh", "th")
    assert did == 1 and s.get_compression_stats()["compressed_docs"] == 1
    assert s.update_document("https://a/b", text="delta epsilon", text_hash="th2", etag='"v2"', stale_count=1)
    assert s.search("delta")[0].doc_id == did and s.search("alpha") == []
    cand = s.get_recrawl_candidates()
    assert cand[0]["etag"] == '"v2"' and cand[0]["stale_count"] == 1
    assert not s.update_document("https://missing", title="x")
    assert s.soft_delete("https://a/b") and s.get_stats()["document_count"] == 0
    with pytest.raises(ValueError):
        LocalStore(tokenizer="evil'); DROP TABLE documents;--")
    s.close()


def test_store_listener_and_export(store):
    events = []
    store.add_listener(lambda ev, p: events.append((ev, p["doc_id"])))
    did = store.add_document("https://x/1", "t", "some words here", "a", "b")
    store.delete_document(did)
    assert events == [("add", did), ("delete", did)]
    store.add_document("https://x/2", "t2", "more words", "a2", "b2")
    assert [d["url"] for d in store.export_documents()] == ["https://x/2"]
    assert store.get_documents_for_publish(limit=5)[0]["title"] == "t2"


def test_ranking_formulas():
    now = 1_000_000.0
    assert freshness_score(now, now=now) == 1.0
    assert freshness_score(now - 7 * 86400, now=now) == pytest.approx(0.5)
    assert freshness_score(now - 365 * 86400, now=now) == 0.05
    assert normalize_bm25(3.0, max_score=3.0) == 0.5 and normalize_bm25(0, max_score=3.0) == 0.0
    assert combined_score(1, 1, 1, 1, title_match=1, url_path=1) == pytest.approx(1.0)
    cands = [RawCandidate(1, "u1", "t", "s", 2.0, now), RawCandidate(2, "u2", "t", "s", 4.0, now - 30 * 86400)]
    ranked = rank_results(cands, limit=1, now=now)
    assert len(ranked) == 1 and ranked[0].doc_id in (1, 2)
    assert rank_results([], limit=5) == []


def test_passage_split_score_select():
    text = ("Intro paragraph that is long enough to be kept as its own passage here.\n\n"
            "short\n\n" + "Sentence one is about GPUs. " * 30 + "\n\n" + "word " * 300)
    ps = passage.split_passages(text, max_length=200)
    assert all(len(p) <= 260 for p in ps) and any("short" in p for p in ps)
    assert passage.score_passage("the quick brown fox", ["quick", "cat"]) == pytest.approx(0.5 + 0.1 * 0.25)
    assert passage.score_passage("", ["x"]) == 0.0
    best = passage.select_best_passage("Nothing relevant here at all, truly.\n\nGPUs run tensor cores fast and "
                                       "tensor memory holds accumulators.", "tensor cores")
    assert "tensor cores" in best
    assert passage.select_best_passage("abc " * 100, "zzz", fallback_length=10) == ("abc " * 100)[:10]
    assert passage.highlight_terms("Tensor cores", ["tensor"]) == "<b>Tensor</b> cores"
    assert passage.title_match_score("Python asyncio guide", ["asyncio", "rust"]) == 0.5
    assert passage.url_path_score("https://x.dev/docs/hooks/", ["react", "hooks"]) == 0.5
    assert passage.classify_intent("download python") == "transactional"
    assert passage.classify_intent("github.com") == "navigational"
    assert passage.classify_intent("what is bm25") == "informational"


def test_nlp_helpers():
    assert "the" in nlp.get_stop_words("en") and nlp.get_stop_words("xx") is nlp.get_stop_words("en")
    assert len(nlp.STOP_WORDS) == 15
    assert nlp.remove_stop_words(["the", "GPU", "is", "fast"]) == ["GPU", "fast"]
    assert nlp.expand_query("database error")[:2] == ["db", "datastore"]
    assert nlp.edit_distance("kitten", "sitting") == 3
    assert nlp.did_you_mean("pyhton asyncio", ["python", "asyncio"]) == ["python asyncio"]
    p = nlp.parse_natural_query("rust tutorials last 3 days site:example.com in korean")
    assert p.cleaned_query == "rust tutorials" and p.include_domains == ["example.com"] and p.language == "ko"
    assert p.date_from == pytest.approx(time.time() - 3 * 86400, abs=5)
    tr = nlp.RelatedSearchTracker()
    tr.record("python asyncio"); tr.record("python asyncio"); tr.record("python gil")
    assert tr.related("python")[0] == "asyncio"


def test_cjk_helpers():
    assert cjk.is_cjk_text("파이썬 비동기") and not cjk.is_cjk_text("python async")
    assert cjk.cjk_bigrams("검색엔진 GPU") == ["검색", "색엔", "엔진", "GPU"]
    assert cjk.cjk_trigrams("ab東京都") == ["ab", "東京都"]
    assert cjk.recommend_tokenizer("東京都の天気") == "trigram" and cjk.recommend_tokenizer("hello") == "unicode61"
    assert cjk.tokenize_query_cjk("hello world") == "hello world"
    assert cjk.tokenize_query_cjk("東京都") == "東京 京都"
    assert cjk.segment_korean("대한민국만세 abc") == ["대한", "한민", "민국", "국만", "만세", "abc"]


def test_sanitize_fts_query():
    assert _sanitize_fts_query('python AND "asyncio" OR (rust)') == "python asyncio rust"
    assert _sanitize_fts_query('"()*') == "infomesh"
    assert len(_sanitize_fts_query("a" * 5000)) == 1000


def test_search_local_ranked_and_snippets(filled_store):
    res = search_local(filled_store, "asyncio python", limit=3)
    assert res.source == "local" and res.total >= 1 and res.results[0].url.endswith("asyncio.html")
    r0 = res.results[0]
    assert 0 < r0.combined_score <= 1.2 and r0.title_match_score > 0 and r0.freshness_score > 0.9
    text = format_fts_results(res)
    assert text.startswith("Found ") and "[1] asyncio" in text and "Score:" in text and "BM25=" in text
    data = json.loads(format_fts_results_json(res))
    assert data["source"] == "local" and data["results"][0]["domain"] == "docs.python.org"
    assert set(data["results"][0]["scores"]) == {"bm25", "freshness", "trust", "authority", "title_match", "url_path"}
    assert format_fts_results(search_local(filled_store, "zzzzqqq")) == "No results found."
    # synonym expansion widens a sparse result set
    assert search_local(filled_store, "db", limit=5).total >= 1


class _FakeVec:
    def __init__(self, rows):
        self.rows = rows

    def add_document(self, **kw):
        pass

    def search(self, query, *, limit=10, min_score=0.0):
        return self.rows[:limit]


def test_hybrid_rrf_merge(filled_store):
    from types import SimpleNamespace as NS

    vec = _FakeVec([NS(doc_id="9", url="https://vec.only/", title="Vec", text_preview="semantic hit", score=0.91),
                    NS(doc_id="1", url="https://docs.python.org/3/library/asyncio.html", title="asyncio",
                       text_preview="p", score=0.8)])
    h = search_hybrid(filled_store, vec, "asyncio", limit=5)
    assert h.source == "hybrid" and h.results[0].source == "hybrid"
    assert h.results[0].combined_score == pytest.approx(1 / 61 + 1 / 62, abs=1e-6)
    assert "RRF=" in format_hybrid_results(h) and "[hybrid]" in format_hybrid_results(h)
    with pytest.raises(TypeError):
        search_hybrid(filled_store, object(), "x")
    assert merge_results([], [], limit=3) == []


def test_query_cache_lru_ttl():
    c = QueryCache(max_size=2, ttl_seconds=10)
    k1, k2, k3 = (QueryCache.make_key(q, 10) for q in ("A", "b", "c"))
    assert k1 == QueryCache.make_key(" a ", 10) and len(k1) == 16
    c.put(k1, 1, now=0); c.put(k2, 2, now=0)
    assert c.get(k1, now=1) == 1
    c.put(k3, 3, now=1)                      # evicts k2 (LRU)
    assert c.get(k2, now=1) is None and c.get(k1, now=20) is None   # TTL
    st = c.stats
    assert st.evictions == 2 and st.hits == 1 and st.misses == 2      # LRU eviction + TTL expiry


def test_format_fetch_result():
    out = format_fetch_result(title="T", url="https://a.b/c", text="body", is_cached=True,
                              crawled_at=time.time() - 10 * 86400)
    assert "is_cached: true" in out and "stale_warning" in out and out.endswith("\n\nbody")
    out = format_fetch_result(title="T", url="https://a.b/c", text="body", is_cached=False, is_paywall=True)
    assert "paywall_warning" in out and "freshly crawled" in out


def _rr(url, title, snippet, score, crawled_at=None, **kw):
    import time as _t

    from infomesh_b200.index.ranking import RankedResult
    return RankedResult(doc_id=url, url=url, title=title, snippet=snippet, bm25_score=kw.get("bm25", score), freshness_score=kw.get("fresh", 0.9),
                        trust_score=0.5, authority_score=0.1, combined_score=score, crawled_at=crawled_at or _t.time())


def test_reranker_parsing_and_llm_fallbacks():
    import asyncio

    from infomesh_b200.search import reranker as RR
    from infomesh_b200.summarizer.engine import LLMBackend, LLMRuntime, ModelInfo

    assert RR._parse_ranking_response("sure: [3, 1, 3, 9, 2]", 4) == [2, 0, 1, 3]
    assert RR._parse_ranking_response("[]", 3) == [0, 1, 2] and RR._parse_ranking_response("no array", 3) is None
    res = [_rr(f"https://e/{i}", f"T{i}", f"snippet {i}", 1.0 - i * 0.1) for i in range(4)]

    class Fixed(LLMBackend):
        def __init__(self, reply, up=True):
            self.reply, self.up = reply, up

        async def generate(self, prompt, *, max_tokens=512):
            assert "1. [T0] snippet 0" in prompt
            if self.reply is None:
                raise RuntimeError("down")
            return self.reply

        async def is_available(self):
            return self.up

        async def model_info(self):
            return ModelInfo("m", LLMRuntime.OLLAMA, None, None, self.up)

    out = asyncio.run(RR.rerank_with_llm("q", res, Fixed("[4, 2]"), max_candidates=3))
    assert [r.title for r in out] == ["T1", "T0", "T2", "T3"]          # 4 is out of range for 3 candidates; tail kept
    assert asyncio.run(RR.rerank_with_llm("q", res, Fixed(None))) == res
    assert asyncio.run(RR.rerank_with_llm("q", res, Fixed("[2,1]", up=False))) == res
    assert asyncio.run(RR.rerank_with_llm("q", res, object())) == res
    assert [r.title for r in asyncio.run(RR.rerank_with_llm("q", res, Fixed("[2, 1]"), top_n=1))] == ["T1"]


def test_explain_facets_quality_crossval_rag_extended():
    import asyncio
    import time

    from infomesh_b200.search import cross_validate as CV
    from infomesh_b200.search import explain as EX
    from infomesh_b200.search import extended as XT
    from infomesh_b200.search import facets as FA
    from infomesh_b200.search import quality as Q
    from infomesh_b200.search import rag as RG

    now = time.time()
    rs = [_rr("https://docs.python.org/a", "Python asyncio guide", "asyncio event loop tutorial for Python tasks", 0.9, now - 3600, bm25=0.95),
          _rr("https://docs.python.org/a/", "Python asyncio guide", "asyncio event loop tutorial for Python tasks", 0.8, now - 3600),
          _rr("https://blog.example/b", "Rust ownership", "ownership borrowing lifetimes explained with examples", 0.7, now - 40 * 86400, fresh=0.1),
          _rr("https://news.example/c", "Asyncio news", "asyncio event loop changes announced in Python 3.13", 0.6, now - 400 * 86400)]
    ex = EX.explain_query("asyncio", "asyncio", rs, 1.25).to_dict()
    assert ex["results"][0]["weighted_contributions"]["bm25"] == round(0.95 * 0.4, 4) and "Strong keyword match" in ex["results"][0]["notes"]
    assert "Stale content — may need recrawl" in ex["results"][2]["notes"] and ex["pipeline"][0] == "sanitize_fts_query"
    f = FA.compute_facets(rs)
    assert f.domains["docs.python.org"] == 2 and f.date_ranges == {"today": 2, "this_year": 1, "older": 1}
    assert [r.url for r in FA.dedup_results(rs)] == [rs[0].url, rs[2].url, rs[3].url]
    assert FA.highlight_snippet("The asyncio Event loop", "event ASYNCIO") == "The **asyncio** **Event** loop"
    cl = FA.cluster_results(rs)
    assert cl and cl[0].label in ("asyncio", "python", "event", "loop") and len(cl[0].results) >= 2
    assert Q.ndcg_at_k([3, 2, 1]) == 1.0 and Q.ndcg_at_k([1, 2, 3]) < 1.0 and Q.mrr([1, 2, 0]) == 0.5
    ab = Q.ABTest("t")
    assert ab.compare("q", [1, 2, 3], [3, 2, 1]).winner == "B" and ab.summary()["B_wins"] == 1
    assert Q.detect_domain_category("https://docs.python.org/3") == "tech-docs" and Q.get_profile("news").freshness_weight == 0.45
    assert Q.extract_temporal_hint("rust news last 3 days") == 3 and Q.extract_temporal_hint("latest gpu") == 7 and Q.extract_temporal_hint("gpu") is None
    ic = Q.QueryIntentClassifier()
    assert ic.classify("how to install docker") == "how_to" and ic.classify("tensor memory") == "informational"
    assert ic.classify_with_confidence("python error exception traceback")[0] == "error_debug"
    dv = Q.diversify_results([{"url": f"https://a.com/{i}"} for i in range(4)] + [{"url": "https://b.com/1"}], max_per_domain=2)
    assert [d["url"] for d in dv] == ["https://a.com/0", "https://b.com/1", "https://a.com/1"]
    P = CV.PeerResult
    rep = CV.cross_validate_results("q", {"p1": [P("p1", "u1", "t", "alpha beta gamma", 1.0), P("p1", "fake", "t", "x", 9.0)],
                                          "p2": [P("p2", "u1", "t", "alpha beta gamma delta", 1.1)],
                                          "p3": [P("p3", "u1", "t", "alpha beta", 0.9)]})
    by = {v.url: v for v in rep.results}
    assert by["u1"].verdict == CV.VERDICT_TRUSTED and by["fake"].verdict == CV.VERDICT_FABRICATED and rep.fabricated_count == 1
    assert CV.cross_validate_results("q", {"p1": [P("p1", "u", "t", "s", 1.0)]}).results[0].verdict == CV.VERDICT_UNVERIFIED
    rag = RG.format_rag_output("asyncio", rs, chunk_size=20, max_chunks=3)
    assert len(rag.chunks) == 3 and rag.chunks[1].chunk_index == 1 and "[Source: Python asyncio guide" in rag.context_window
    ans = RG.extract_answers("asyncio event loop", rs)
    assert ans and ans[0].confidence > 0.5 and "Search Results:" in RG.build_summary_prompt("q", rs)
    ents = RG.extract_entities("Python and Docker with Python by Guido Van Rossum", source_url="u")
    assert ents[0].text == "Python" and ents[0].count == 2 and any(e.entity_type == "NAME" for e in ents)
    assert RG.compute_toxicity_score("scam phishing malware here") > 0.3 and len(RG.filter_by_toxicity(rs)) == 4
    assert '"index" and "score"' in RG.build_cot_rerank_prompt("q", rs)

    async def fn(q, k, lang):
        if q == "bad":
            raise ValueError("boom")
        return [{"q": q, "k": k}]

    br = asyncio.run(XT.batch_search([XT.BatchQuery("a", 2), XT.BatchQuery("bad")], fn))
    assert br.total_queries == 2 and br.results[0].results == [{"q": "a", "k": 2}] and br.results[1].error == "boom"
    sc = XT.SummaryCache(max_entries=2, ttl_seconds=100)
    sc.put("Q one", "s1", ["u"])
    sc.put("q two", "s2", [])
    sc.put("q three", "s3", [])
    assert sc.size == 2 and sc.get("q ONE ") is None and sc.get("q three").summary == "s3"
    assert XT.translate_query_keywords("파이썬 설치 오류", "ko") == ["install", "error"] and XT.translate_query_keywords("x", "xx") == []
