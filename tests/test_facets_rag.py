"""search/facets.py + search/rag.py — facets, clustering, highlighting, near-duplicate removal, RAG formatting, answer
extraction, entities, toxicity."""
import time

from infomesh_b200.index.ranking import RankedResult
from infomesh_b200.search import facets as F
from infomesh_b200.search import rag as R


def rr(url, title="T", snippet="s", score=0.5, age_days=0.5):
    return RankedResult(doc_id=url, url=url, title=title, snippet=snippet, bm25_score=score, freshness_score=0.5, trust_score=0.5,
                        authority_score=0.1, combined_score=score, crawled_at=time.time() - age_days * 86400)


def test_compute_facets_domains_dates_languages():
    rs = [rr("https://a.com/1"), rr("https://a.com/2", age_days=3), rr("https://b.org/x", age_days=100),
          rr("https://c.io/y", age_days=800)]
    f = F.compute_facets(rs, languages={"https://a.com/1": "en", "https://b.org/x": "ko"})
    assert f.domains == {"a.com": 2, "b.org": 1, "c.io": 1}
    assert f.languages == {"en": 1, "ko": 1}
    assert f.date_ranges == {"today": 1, "this_week": 1, "this_year": 1, "older": 1}
    assert list(f.to_dict()["domains"])[0] == "a.com"


def test_age_bucket_unknown_for_missing_or_future_timestamps():
    now = time.time()
    assert F._age_bucket(0, now) == "unknown" and F._age_bucket(now + 3 * 86400, now) == "unknown"
    assert F._age_bucket(now - 20 * 86400, now) == "this_month"


def test_cluster_results_groups_by_shared_keywords():
    rs = [rr("u1", "Python asyncio tutorial", "event loop", 0.9), rr("u2", "Asyncio patterns", "python tasks", 0.7),
          rr("u3", "Rust ownership", "borrow checker", 0.5), rr("u4", "Rust lifetimes", "borrow rules", 0.4)]
    cl = F.cluster_results(rs)
    labels = {c.label: [r.url for r in c.results] for c in cl}
    assert any(set(v) == {"u1", "u2"} for v in labels.values()) and any(set(v) == {"u3", "u4"} for v in labels.values())
    assert cl[0].score >= cl[-1].score and F.cluster_results(rs[:1]) == []


def test_highlight_snippet_and_custom_marker():
    assert F.highlight_snippet("Python loves python", "python") == "**Python** loves **python**"
    assert F.highlight_snippet("abc", "", marker="__") == "abc"
    assert F.highlight_snippet("fast api", "api", marker="<<") == "fast <<api<<"


def test_dedup_results_by_url_and_text_similarity():
    a = rr("https://a.com/p", "Python asyncio guide", "learn the event loop today")
    b = rr("https://a.com/p/", "different title", "different words entirely here")
    c = rr("https://mirror.com/p", "Python asyncio guide", "learn the event loop today")
    d = rr("https://d.com/", "Rust ownership", "borrow checker explained")
    assert [r.url for r in F.dedup_results([a, b, c, d])] == ["https://a.com/p", "https://d.com/"]


def test_format_rag_output_chunks_and_context_window():
    long = "x" * 1200
    out = R.format_rag_output("q", [rr("u1", "Doc1", long, 0.9), rr("u2", "Doc2", "short text", 0.4)], chunk_size=500)
    assert [c.chunk_index for c in out.chunks] == [0, 1, 2, 0] and out.total_results == 2
    assert out.chunks[3].metadata["bm25_score"] == 0.4 and out.chunks[0].metadata == {}
    assert "[Source: Doc2 (u2)]\nshort text" in out.context_window
    assert R.format_rag_output("q", [rr("u", snippet="x" * 5000)], chunk_size=100, max_chunks=3).to_dict()["chunks"][2]["chunk_index"] == 2


def test_extract_answers_prefers_covering_sentences_and_strips_markup():
    rs = [rr("u1", "Doc", "The <b>event</b> loop schedules coroutines. Unrelated trivia about cats here.", 0.9),
          rr("u2", "Doc2", "Nothing relevant in this one at all.", 0.1)]
    ans = R.extract_answers("event loop coroutines", rs)
    assert ans and ans[0].answer == "The event loop schedules coroutines" and ans[0].source_url == "u1"
    assert all("<b>" not in a.answer for a in ans) and 0 < ans[0].confidence <= 1


def test_build_prompts_respect_budgets():
    rs = [rr(f"u{i}", f"Title{i}", "s" * 400) for i in range(20)]
    p = R.build_summary_prompt("my query", rs, max_context=1000)
    assert '"my query"' in p and "[1] Title0" in p and "[4] Title3" not in p
    cot = R.build_cot_rerank_prompt("q", rs, max_candidates=3)
    assert "3. [Title2]" in cot and "4. [Title3]" not in cot and "JSON array" in cot


def test_extract_entities_counts_tech_and_names():
    text = "Guido van Rossum created Python. Python and Rust run on Linux. Ada Lovelace wrote notes."
    ents = {(e.entity_type, e.text): e for e in R.extract_entities(text, source_url="https://x")}
    assert ents[("TECH", "Python")].count == 2 and ents[("TECH", "Rust")].source_urls == ["https://x"]
    assert ("NAME", "Ada Lovelace") in ents


def test_toxicity_score_and_filter():
    assert R.compute_toxicity_score("") == 0.0
    assert R.compute_toxicity_score("a friendly tutorial about python") == 0.0
    assert R.compute_toxicity_score("scam phishing malware") == 1.0
    rs = [rr("ok", snippet="a clean page about gardening and soil"), rr("bad", snippet="scam phishing malware scam")]
    assert [r.url for r in F.dedup_results(R.filter_by_toxicity(rs))] == ["ok"]
