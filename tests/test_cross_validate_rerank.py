"""search/cross_validate.py, search/reranker.py (LLM path with a fake backend), search/extended.py."""
import asyncio
import time

from infomesh_b200.index.ranking import RankedResult
from infomesh_b200.search import cross_validate as CV
from infomesh_b200.search import extended as X
from infomesh_b200.search import reranker as RR
from infomesh_b200.summarizer.engine import LLMBackend, ModelInfo


def pr(peer, url, score=1.0, snippet="python asyncio event loop guide"):
    return CV.PeerResult(peer, url, "T", snippet, score)


def test_cross_validation_skipped_with_one_peer():
    rep = CV.cross_validate_results("q", {"p1": [pr("p1", "u1"), pr("p1", "u1")]})
    assert rep.total_peers == 1 and [v.verdict for v in rep.results] == [CV.VERDICT_UNVERIFIED]
    assert "insufficient" in rep.detail


def test_cross_validation_trusted_suspicious_fabricated():
    peers = {
        "p1": [pr("p1", "agree"), pr("p1", "dev", 1.0), pr("p1", "snip"), pr("p1", "lonely")],
        "p2": [pr("p2", "agree"), pr("p2", "dev", 1.0), pr("p2", "snip", snippet="completely different words about cooking pasta")],
        "p3": [pr("p3", "agree"), pr("p3", "dev", 50.0)],
    }
    rep = CV.cross_validate_results("q", peers)
    v = {r.url: r for r in rep.results}
    assert v["agree"].verdict == CV.VERDICT_TRUSTED and v["agree"].agreement_ratio == 1.0
    assert v["dev"].verdict == CV.VERDICT_SUSPICIOUS and "deviation" in v["dev"].detail
    assert v["snip"].verdict == CV.VERDICT_SUSPICIOUS and "snippet" in v["snip"].detail
    assert v["lonely"].verdict == CV.VERDICT_FABRICATED and v["lonely"].appearing_peers == ["p1"]
    assert rep.suspicious_count == 2 and rep.fabricated_count == 1 and rep.results[0].url == "agree"


def test_one_vote_per_peer_and_snippet_similarity():
    rep = CV.cross_validate_results("q", {"p1": [pr("p1", "u"), pr("p1", "u")], "p2": [pr("p2", "other")]})
    assert {r.url: len(r.appearing_peers) for r in rep.results} == {"u": 1, "other": 1}
    assert CV.snippet_similarity("a b c", "a b d") == 0.5 and CV.snippet_similarity("", "x") == 0.0


def rr(url, title, score):
    return RankedResult(url, url, title, f"snippet of {title}", score, 0.5, 0.5, 0.1, score, time.time())


class FakeLLM(LLMBackend):
    def __init__(self, reply="[3, 1, 2]", available=True, boom=False):
        self.reply, self.available, self.boom, self.prompts = reply, available, boom, []

    async def generate(self, prompt, *, max_tokens=512):
        if self.boom:
            raise RuntimeError("backend down")
        self.prompts.append(prompt)
        return self.reply

    async def is_available(self):
        return self.available

    async def model_info(self):
        return ModelInfo("fake", None, None, None, True)  # type: ignore[arg-type]


def test_parse_ranking_response_tolerates_junk():
    assert RR._parse_ranking_response("Here: [2, 2, 9, 1] done", 3) == [1, 0, 2]      # repeats / out-of-range dropped
    assert RR._parse_ranking_response('[2, "x", 1]', 3) is None                       # only digit arrays are recognised
    assert RR._parse_ranking_response("no array here", 3) is None
    assert RR._parse_ranking_response("[1, 2", 3) is None


def test_rerank_with_llm_applies_permutation_and_keeps_tail():
    rs = [rr("a", "A", 0.9), rr("b", "B", 0.8), rr("c", "C", 0.7), rr("d", "D", 0.6)]
    llm = FakeLLM("Ranking: [3, 1, 2]")
    out = asyncio.run(RR.rerank_with_llm("q", rs, llm, max_candidates=3))
    assert [r.url for r in out] == ["c", "a", "b", "d"] and "1. [A]" in llm.prompts[0]
    assert [r.url for r in asyncio.run(RR.rerank_with_llm("q", rs, llm, top_n=2, max_candidates=3))] == ["c", "a"]


def test_rerank_with_llm_degrades_to_first_stage_order():
    rs = [rr("a", "A", 0.9), rr("b", "B", 0.8)]
    for llm in (FakeLLM(available=False), FakeLLM(boom=True), FakeLLM("garbage"), object()):
        assert asyncio.run(RR.rerank_with_llm("q", rs, llm)) == rs
    assert asyncio.run(RR.rerank_with_llm("q", [], FakeLLM())) == []


def test_batch_search_isolates_failures_and_limits_parallelism():
    running = peak = 0

    async def fn(q, k, lang):
        nonlocal running, peak
        running += 1
        peak = max(peak, running)
        await asyncio.sleep(0.01)
        running -= 1
        if q == "bad":
            raise ValueError("nope")
        return [{"url": f"https://x/{q}", "k": k}]

    qs = [X.BatchQuery(f"q{i}", top_k=3) for i in range(6)] + [X.BatchQuery("bad")]
    resp = asyncio.run(X.batch_search(qs, fn, max_parallel=2))
    assert resp.total_queries == 7 and peak <= 2
    assert resp.results[0].results == [{"url": "https://x/q0", "k": 3}] and resp.results[-1].error == "nope"


def test_summary_cache_ttl_and_eviction():
    c = X.SummaryCache(max_entries=2, ttl_seconds=0.05)
    c.put("Hello World", "s1", ["u"])
    assert c.get("  hello world ").summary == "s1" and c.size == 1
    c.put("b", "s2", [])
    c.put("c", "s3", [])                      # evicts the oldest
    assert c.size == 2 and c.get("hello world") is None
    time.sleep(0.06)
    assert c.get("b") is None and c.size == 1


def test_translate_query_keywords():
    assert X.translate_query_keywords("파이썬 설치 오류", "ko") == ["install", "error"]
    assert X.translate_query_keywords("安装 数据库", "zh") == ["install", "database"]
    assert X.translate_query_keywords("hello", "xx") == []
