"""observability/metrics.py, data_quality.py, summarizer/verify.py, utils/tokenizer.py."""
import threading
import time
from types import SimpleNamespace

from infomesh_b200 import data_quality as DQ
from infomesh_b200.observability import metrics as M
from infomesh_b200.summarizer import verify as V
from infomesh_b200.utils import tokenizer as T


# ------------------------------------------------------------------ metrics
def test_collector_counters_gauges_summaries_and_prometheus_text():
    c = M.MetricsCollector()
    c.inc("search.total")
    c.inc("search.total", 2)
    c.set_gauge("peers-connected", 7)
    for v in (10, 20, 30, 40):
        c.observe("latency ms", v)
    text = c.format_prometheus()
    assert "# TYPE search_total counter\nsearch_total 3.0" in text and "peers_connected 7" in text
    assert 'latency_ms{quantile="0.5"} 30.000' in text and "latency_ms_count 4" in text and "latency_ms_avg 25.000" in text
    assert text.rstrip().splitlines()[-1].startswith("infomesh_uptime_seconds")
    d = c.to_dict()
    assert d["counters"]["search.total"] == 3 and d["histograms"]["latency ms"]["sum"] == 100


def test_collector_timer_and_window_and_threads():
    c = M.MetricsCollector()
    with c.timer("t"):
        time.sleep(0.01)
    assert c.to_dict()["histograms"]["t"]["avg"] >= 5
    for i in range(1500):
        c.observe("w", i)
    assert c.to_dict()["histograms"]["w"]["count"] == 1000              # sliding window
    ths = [threading.Thread(target=lambda: [c.inc("n") for _ in range(500)]) for _ in range(8)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert c.to_dict()["counters"]["n"] == 4000
    assert M.get_collector() is M.get_collector()


def test_query_trace_benchmark_dashboards():
    tr = M.QueryTrace("t1", "q")
    tr.add_span(M.QuerySpan("s1", "p1", "local", latency_ms=3.0))
    tr.add_span(M.QuerySpan("s2", "p2", "remote", latency_ms=7.5, metadata={"hops": "2"}))
    assert tr.total_latency_ms == 10.5 and tr.to_dict()["spans"][1]["metadata"] == {"hops": "2"}
    b = M.run_benchmark("noop", lambda: None, iterations=20)
    assert b.iterations == 20 and b.min_ms <= b.p50_ms <= b.p99_ms <= b.max_ms
    dash = M.generate_grafana_dashboard()["dashboard"]
    assert len(dash["panels"]) >= 6 and any("gpu" in p["targets"][0]["expr"] for p in dash["panels"])
    rules = M.generate_alert_rules()
    assert {r["alert"] for r in rules} >= {"HighSearchLatency", "NoPeersConnected"} and all("expr" in r for r in rules)
    assert M.configure_log_forwarding(format="console")["processors"] == "ConsoleRenderer()"


# ------------------------------------------------------------------ data quality
def test_freshness_indicator_labels_and_grades():
    now = 10_000_000.0
    cases = [(30, "just now", "A"), (150, "2 minutes ago", "A"), (3600, "1 hour ago", "A"), (3 * 86400, "3 days ago", "B"),
             (15 * 86400, "2 weeks ago", "C"), (60 * 86400, "2 months ago", "D"), (400 * 86400, "13 months ago", "F")]
    for age, label, grade in cases:
        f = DQ.compute_freshness_indicator(now - age, now=now)
        assert (f.age_label, f.freshness_grade) == (label, grade), age


def test_trust_grades():
    assert [DQ.compute_trust_grade(s).grade for s in (0.95, 0.85, 0.7, 0.55, 0.35, 0.1)] == ["A+", "A", "B", "C", "D", "F"]
    assert DQ.compute_trust_grade(0.1).color == "red"


def test_extract_citations_kinds_and_dedup():
    text = ("See doi 10.1000/xyz123 and arXiv:2101.12345v2, RFC 2616, ISBN 978-3-16-148410-0, "
            "https://example.com/a?b=1 and again https://example.com/a?b=1 .")
    kinds = {c.citation_type: c.identifier for c in DQ.extract_citations(text)}
    assert kinds["doi"] == "10.1000/xyz123" and kinds["rfc"] == "RFC 2616" and kinds["url"].startswith("https://example.com/a")
    assert "arxiv" in kinds and "isbn" in kinds
    assert sum(c.citation_type == "url" for c in DQ.extract_citations(text)) == 1


def test_cross_reference_results_verdicts():
    r = lambda s, u: SimpleNamespace(snippet=s, url=u)  # noqa: E731
    claim = "python asyncio event loop"
    sup = DQ.cross_reference_results(claim, [r("the <b>asyncio</b> event loop in python", "a"), r("python event loop", "b"), r("cats", "c")])
    assert sup.verdict == "supported" and sup.supporting_sources == 2 and sup.sources == ["a", "b"]
    dis = DQ.cross_reference_results(claim, [r("asyncio event loop", "a"), r("cats", "b"), r("dogs", "c")])
    assert dis.verdict == "disputed" and dis.contradicting_sources == 2
    assert DQ.cross_reference_results(claim, [r("cats", "a")]).verdict == "unverified"
    assert DQ.cross_reference_results("", []).verdict == "unverified" and DQ.cross_reference_results(claim, []).confidence == 0.0


# ------------------------------------------------------------------ summary verification
SRC = ("InfoMesh was released in 2024 by the Open Search Collective. It indexes 10 million documents across 8 nodes. "
       "The crawler respects robots.txt and waits 2 seconds between requests. About 95% of queries finish within 50 milliseconds. "
       "It is fun.")


def test_key_facts_prefer_numbers_and_names():
    facts = V.extract_key_facts(SRC, max_facts=3)
    assert len(facts) == 3 and all(any(ch.isdigit() for ch in f.text) for f in facts)
    assert SRC[facts[0].source_offset:].lstrip().startswith(facts[0].text[:10])


def test_self_verify_pass_low_coverage_and_contradiction():
    good = "InfoMesh, released in 2024 by the Open Search Collective, indexes 10 million documents across 8 nodes; 95% of queries finish within 50 milliseconds."
    assert V.self_verify(SRC, good).passed
    bad = V.self_verify(SRC, "A short note about gardening and tomatoes.")
    assert not bad.passed and "low coverage" in bad.detail
    lie = V.self_verify(SRC, good + " It cost 777 dollars, has 31337 users and 4242 servers.")
    assert lie.has_contradiction and not lie.passed


def test_cross_validate_and_report_levels():
    ours = "InfoMesh indexes 10 million documents across 8 nodes in 2024 by the Open Search Collective with 95% of queries within 50 milliseconds"
    assert V.cross_validate(ours, []).passed is False
    assert V.cross_validate(ours, [ours, ours + " quickly"]).passed
    assert V.compute_similarity("a b", "") == 0.0
    rep = V.verify_summary("u", "h", SRC, ours, peer_summaries=[ours])
    assert rep.level == V.VerificationLevel.CROSS_VALIDATED and rep.quality_score > 0.8
    solo = V.verify_summary("u", "h", SRC, ours)
    assert solo.level == V.VerificationLevel.SELF_VERIFIED and solo.cross_check is None
    assert V.verify_summary("u", "h", SRC, "tomatoes").level == V.VerificationLevel.UNVERIFIED


# ------------------------------------------------------------------ tokenizers
def test_fnv1a_reference_vectors():
    assert T.fnv1a(b"") == 0xCBF29CE484222325 and T.fnv1a(b"a") == 0xAF63DC4C8601EC8C


def test_hash_tokenizer_is_deterministic_and_bounded():
    tok = T.HashTokenizer(30522)
    ids = tok.encode("Hello, hello WORLD!")
    assert ids[0] == 101 and ids[-1] == 102 and ids[1] == ids[2] and all(1000 <= i < 30522 for i in ids[1:-1])
    assert len(tok.encode("w " * 600, max_len=64)) == 64 and tok.encode_plain("a b c", 2) == ids_plain(tok, "a b")
    batch, lens = tok.encode_batch(["one two three", "one"], max_len=32, pad_to_multiple=8)
    assert tuple(batch.shape) == (2, 8) and lens.tolist() == [5, 3] and batch[1, 3:].tolist() == [0] * 5
    assert tok.decode([101, 5]) == "<101> <5>"


def ids_plain(tok, text):
    return [tok.word_id(w) for w in tok.words(text)]


def test_hash_tokenizer_with_vocab_file(tmp_path):
    vf = tmp_path / "vocab.txt"
    vf.write_text("[PAD]\nhello\nworld\n", encoding="utf-8")
    tok = T.HashTokenizer(100, vocab_file=str(vf))
    assert tok.encode("hello there world", add_special=False) == [1, tok.sp.unk, 2]
    assert tok.decode([1, 2, 99]) == "hello world [UNK]"


def test_seq2seq_tokenizer_eos_and_decode_roundtrip():
    tok = T.load_tokenizer("t5-small")
    ids = tok.encode("summarize this text please")
    assert ids[-1] == 1 and len(ids) == 5
    assert tok.decode(ids + [0, 0, 7]) == "summarize this text please"
    assert len(tok.encode("w " * 100, max_len=10)) == 10 and tok.encode("x", add_eos=False)[-1] != 1
