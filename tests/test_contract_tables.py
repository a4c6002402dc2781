"""Behavioural-contract tables as parametrised cases: one assertion family per row, so a
regression names the exact constant or rule that moved."""
from __future__ import annotations

import hashlib
import math
import time

import pytest

from infomesh_b200.crawler import dedup, freshness, simhash as SH
from infomesh_b200.credits.types import ActionType, ContributionTier
from infomesh_b200.errors import ErrorCategory, format_error, get_error
from infomesh_b200.hashing import content_hash, short_hash
from infomesh_b200.index import ranking as R
from infomesh_b200.index.distributed import extract_keywords
from infomesh_b200.p2p import protocol as P
from infomesh_b200.resources import governor as G
from infomesh_b200.search import cjk, nlp
from infomesh_b200.search.cache import QueryCache
from infomesh_b200.search.query import sanitize_fts_query
from infomesh_b200.security import SSRFError, is_safe_url, validate_url
from infomesh_b200.trust import scoring as TS

DAY = 86400.0


# ------------------------------------------------------------------ ranking (Appendix B row 1-2)
@pytest.mark.parametrize("age_days,expected", [(0, 1.0), (7, 0.5), (14, 0.25), (21, 0.125), (28, 0.0625), (35, 0.05), (365, 0.05),
                                                (-3, 1.0), (3.5, 2 ** -0.5), (1, 2 ** (-1 / 7))])
def test_freshness_halves_every_week_with_floor(age_days, expected):
    now = 1_700_000_000.0
    assert R.freshness_score(now - age_days * DAY, now=now) == pytest.approx(expected, rel=1e-9)


@pytest.mark.parametrize("score,top,expected", [(0, 5, 0.0), (-1, 5, 0.0), (5, 5, 0.5), (1, 3, 0.25), (9, 1, 0.9), (2, 0, 1.0), (1e-9, 1, 1e-9)])
def test_bm25_squash(score, top, expected):
    assert R.normalize_bm25(score, max_score=top) == pytest.approx(expected, rel=1e-6)


@pytest.mark.parametrize("signal,weight", [("bm25", 0.40), ("freshness", 0.15), ("trust", 0.10), ("authority", 0.15), ("title_match", 0.15),
                                            ("url_path", 0.05)])
def test_each_signal_contributes_its_weight(signal, weight):
    kw = {"bm25": 0.0, "freshness": 0.0, "trust": 0.0, "authority": 0.0, "title_match": 0.0, "url_path": 0.0}
    kw[signal] = 1.0
    pos = (kw.pop("bm25"), kw.pop("freshness"), kw.pop("trust"), kw.pop("authority"))
    assert R.combined_score(*pos, **kw) == pytest.approx(weight)


def test_weights_sum_to_one_and_match_the_signal_table():
    assert sum(w for _, w in R.SIGNALS) == pytest.approx(1.0)
    assert R.weight_vector({"w_bm25": 1.0})[0] == 1.0
    with pytest.raises(TypeError):
        R.weight_vector({"w_bogus": 1.0})


@pytest.mark.parametrize("n", [1, 2, 7, 30])
def test_rank_results_is_sorted_stable_and_rounded(n):
    now = 1_700_000_000.0
    cands = [R.RawCandidate(doc_id=i, url=f"u{i}", title="", snippet="", bm25_raw=float((i * 7) % 5 + 1), crawled_at=now - i * DAY) for i in range(n)]
    out = R.rank_results(cands, limit=10, now=now)
    scores = [r.combined_score for r in out]
    assert scores == sorted(scores, reverse=True) and len(out) == min(n, 10)
    assert all(round(r.combined_score, 6) == r.combined_score and 0 < r.bm25_score <= 0.5 + 1e-9 for r in out)
    again = R.rank_results(cands, limit=10, now=now)
    assert [r.doc_id for r in again] == [r.doc_id for r in out]


# ------------------------------------------------------------------ freshness tiers / recrawl classes
@pytest.mark.parametrize("age,tier", [(0, "hot"), (3599, "hot"), (3600, "hot"), (3601, "warm"), (86400, "warm"), (86401, "cold"),
                                      (604800, "cold"), (604801, "stale"), (10 ** 8, "stale")])
def test_freshness_tier_boundaries(age, tier):
    assert freshness.classify_freshness(1000.0, now=1000.0 + age) == tier


@pytest.mark.parametrize("trigger,prio", [("user_request", 0), ("rss_update", 1), ("content_change", 2), ("peer_announce", 3), ("scheduled", 4)])
def test_trigger_classes(trigger, prio):
    assert freshness.TRIGGER_PRIORITY[freshness.RecrawlTrigger(trigger)] == prio


def test_recrawl_queue_serves_all_classes_in_order():
    q = freshness.PriorityRecrawlQueue()
    order = ["scheduled", "peer_announce", "content_change", "rss_update", "user_request"]
    for i, t in enumerate(order):
        assert q.enqueue(f"u-{t}", freshness.RecrawlTrigger(t), now=100 - i)
    assert [q.dequeue().trigger.value for _ in range(5)] == list(reversed(order))


# ------------------------------------------------------------------ credits / trust constants
@pytest.mark.parametrize("action,weight", [(ActionType.CRAWL, 1.0), (ActionType.QUERY_PROCESS, 0.5), (ActionType.DOC_HOSTING, 0.1),
                                            (ActionType.NETWORK_UPTIME, 0.5), (ActionType.LLM_SUMMARIZE_OWN, 1.5), (ActionType.LLM_SUMMARIZE_PEER, 2.0)])
def test_action_weights(action, weight):
    from infomesh_b200.credits import types as CT

    table = next(v for k, v in vars(CT).items() if isinstance(v, dict) and ActionType.CRAWL in v)
    assert table[action] == weight


@pytest.mark.parametrize("score,tier,cost", [(0, ContributionTier.TIER_1, 0.100), (99.9, ContributionTier.TIER_1, 0.100),
                                              (100, ContributionTier.TIER_2, 0.050), (999, ContributionTier.TIER_2, 0.050),
                                              (1000, ContributionTier.TIER_3, 0.033), (10 ** 6, ContributionTier.TIER_3, 0.033)])
def test_contribution_tiers_and_search_cost(score, tier, cost):
    from infomesh_b200.credits.types import TIER_THRESHOLDS

    got = next((t, c) for floor, t, c in TIER_THRESHOLDS if score >= floor)
    assert got == (tier, cost)


@pytest.mark.parametrize("score,tier", [(1.0, "TRUSTED"), (0.8, "TRUSTED"), (0.79, "NORMAL"), (0.5, "NORMAL"), (0.49, "SUSPECT"), (0.3, "SUSPECT"),
                                        (0.29, "UNTRUSTED"), (0.0, "UNTRUSTED")])
def test_trust_tiers(score, tier):
    assert TS.trust_tier(score).name == tier


def test_trust_score_component_weights():
    full = TS.compute_trust_score(uptime_hours=24 * 365, contribution_raw=10 ** 6, audit_total=10, audit_passed=10, summary_avg=1.0)
    none = TS.compute_trust_score(uptime_hours=0, contribution_raw=0, audit_total=10, audit_passed=0, summary_avg=0.0)
    assert full == pytest.approx(1.0) and none < 0.15     # (no summary ratings count as the neutral 0.5 x 20 %)
    passed = TS.compute_trust_score(uptime_hours=0, contribution_raw=0, audit_total=4, audit_passed=4, summary_avg=0.0)
    failed = TS.compute_trust_score(uptime_hours=0, contribution_raw=0, audit_total=4, audit_passed=0, summary_avg=0.0)
    assert passed - failed == pytest.approx(0.40)            # the audit component carries 40 %
    rated = TS.compute_trust_score(uptime_hours=0, contribution_raw=0, audit_total=4, audit_passed=0, summary_avg=1.0)
    assert rated - failed == pytest.approx(0.10)             # summary quality 20 %, measured from its neutral 0.5 default


# ------------------------------------------------------------------ governor ladder
@pytest.mark.parametrize("cpu,mem,ratio,level", [(10, 10, 0.1, "NORMAL"), (61, 10, 0.1, "WARNING"), (10, 71, 0.1, "WARNING"), (10, 10, 0.75, "WARNING"),
                                                  (81, 10, 0.1, "OVERLOADED"), (10, 86, 0.1, "OVERLOADED"), (10, 10, 0.9, "OVERLOADED"),
                                                  (91, 10, 0.1, "SEVERE"), (10, 91, 0.1, "SEVERE"), (10, 10, 1.0, "SEVERE"),
                                                  (96, 10, 0.1, "DEFENSIVE"), (10, 96, 0.1, "DEFENSIVE"), (10, 10, 1.2, "DEFENSIVE")])
def test_governor_classification(cpu, mem, ratio, level):
    assert G.classify(cpu, mem, ratio).name == level


# ------------------------------------------------------------------ URL normalisation / SSRF
@pytest.mark.parametrize("raw,norm", [
    ("HTTP://Example.COM/Path/", "http://example.com/Path"), ("https://e.org/a#frag", "https://e.org/a"),
    ("https://e.org/a?b=2&a=1", "https://e.org/a?a=1&b=2"), ("https://e.org/a?utm_source=x&id=3", "https://e.org/a?id=3"),
    ("https://e.org/a?fbclid=zzz", "https://e.org/a"), ("https://e.org/?gclid=1&utm_medium=m&utm_campaign=c", "https://e.org/"),
    ("https://e.org", "https://e.org/"), ("https://e.org/a/b/../c", "https://e.org/a/b/../c")])
def test_url_normalisation(raw, norm):
    assert dedup.normalize_url(raw) == norm


@pytest.mark.parametrize("url", ["http://127.0.0.1/", "http://localhost/x", "http://10.1.2.3/", "http://192.168.0.5:8080/", "http://172.16.9.9/",
                                  "http://169.254.169.254/latest/meta-data", "http://[::1]/", "ftp://example.com/", "file:///etc/passwd",
                                  "http://0.0.0.0/", "javascript:alert(1)", "http://metadata.google.internal/"])
def test_ssrf_guard_blocks(url):
    assert not is_safe_url(url)
    with pytest.raises(SSRFError):
        validate_url(url)


@pytest.mark.parametrize("url", ["https://example.com/", "http://example.org:8080/a?b=c", "https://sub.domain.example.co.uk/x"])
def test_ssrf_guard_allows_public_hosts(url):
    assert validate_url(url) == url and is_safe_url(url)


# ------------------------------------------------------------------ FTS sanitiser / keywords / NLP / CJK
@pytest.mark.parametrize("raw,clean", [('hello "world"', "hello world"), ("a AND b", "a b"), ("x OR y NOT z", "x y z"), ("foo* (bar)", "foo bar"),
                                       ("col:value", "col value"), ("NEAR(a b)", "a b"), ("", "infomesh"), ('"()*', "infomesh"),
                                       ("   spaced    out  ", "spaced out"), ("and or not", "and or not")])
def test_fts_sanitiser(raw, clean):
    assert sanitize_fts_query(raw) == clean


def test_fts_sanitiser_caps_length():
    assert len(sanitize_fts_query("word " * 1000)) <= 1000


@pytest.mark.parametrize("text,expect_in,expect_out", [
    ("The quick brown fox jumps over the lazy dog", ["quick", "brown", "fox"], ["the", "over"]),
    ("GPU gpu Gpu tensor tensor core", ["gpu", "tensor", "core"], []),
    ("a b c dd ee", ["dd", "ee"], ["a", "b", "c"]),
    ("x86 arm64 and riscv", ["x86", "arm64", "riscv"], ["and"])])
def test_keyword_extraction_rules(text, expect_in, expect_out):
    kws = extract_keywords(text)
    assert all(k in kws for k in expect_in) and not any(k in kws for k in expect_out) and kws == [k.lower() for k in kws]


def test_keyword_extraction_orders_by_frequency_and_caps():
    text = " ".join(f"w{i} " * (60 - i) for i in range(60))
    kws = extract_keywords(text)
    assert len(kws) == 50 and kws[0] == "w0" and "w59" not in kws
    assert extract_keywords(text, max_keywords=5) == ["w0", "w1", "w2", "w3", "w4"]


@pytest.mark.parametrize("lang,stop,keep", [("en", "the", "tensor"), ("ko", "그리고", "텐서"), ("ja", "これ", "検索"), ("zh", "的", "搜索"), ("es", "el", "buscar"),
                                             ("fr", "les", "chercher"), ("de", "und", "suchen"), ("pt", "não", "buscar"), ("hi", "और", "खोज"),
                                             ("ru", "что", "поиск"), ("th", "และ", "ค้นหา"), ("vi", "của", "kiếm"), ("id", "yang", "pencarian"),
                                             ("tr", "ve", "aramak"), ("ar", "في", "بحث")])
def test_stop_words_for_every_language(lang, stop, keep):
    words = nlp.get_stop_words(lang)
    assert stop in words and keep not in words
    assert nlp.remove_stop_words([stop, keep], lang) == [keep]


@pytest.mark.parametrize("a,b,d", [("", "", 0), ("a", "", 1), ("kitten", "sitting", 3), ("flaw", "lawn", 2), ("tensor", "tensor", 0), ("abc", "acb", 2)])
def test_edit_distance(a, b, d):
    assert nlp.edit_distance(a, b) == d and nlp.edit_distance(b, a) == d


@pytest.mark.parametrize("text,is_cjk", [("hello world", False), ("검색 엔진 최적화", True), ("東京都の人口", True), ("搜索引擎", True), ("mixed 한 word only here ok", False), ("", False)])
def test_cjk_detection(text, is_cjk):
    assert cjk.is_cjk_text(text) is is_cjk


@pytest.mark.parametrize("text,grams", [("검색엔진", ["검색", "색엔", "엔진"]), ("東京", ["東京"]), ("搜索引擎优化", ["搜索", "索引", "引擎", "擎优", "优化"])])
def test_cjk_bigrams(text, grams):
    assert cjk.cjk_bigrams(text) == grams


# ------------------------------------------------------------------ SimHash contract
def test_simhash_is_md5_first8_big_endian_majority_vote():
    text = "alpha beta gamma"           # exactly one 3-word shingle -> the fingerprint IS md5(shingle)[:8] big-endian
    want = int.from_bytes(hashlib.md5(b"alpha beta gamma").digest()[:8], "big")
    assert SH.simhash(text) == want
    assert SH.simhash("") == 0 and SH.simhash("one two") == int.from_bytes(hashlib.md5(b"one two").digest()[:8], "big")


@pytest.mark.parametrize("a,b,d", [(0, 0, 0), (0, 1, 1), (0b1011, 0b0001, 2), (2 ** 64 - 1, 0, 64), (0xF0F0, 0x0F0F, 16)])
def test_hamming_distance(a, b, d):
    assert SH.hamming_distance(a, b) == d and SH.is_near_duplicate(a, b) is (d <= 3)


def test_simhash_near_duplicates_and_unrelated_texts():
    base = " ".join(f"word{i}" for i in range(400))
    near = base + " trailing"
    far = " ".join(f"other{i}" for i in range(400))
    assert SH.hamming_distance(SH.simhash(base), SH.simhash(near)) <= 3
    assert SH.hamming_distance(SH.simhash(base), SH.simhash(far)) > 10


# ------------------------------------------------------------------ wire protocol
@pytest.mark.parametrize("mt", list(P.MessageType))
def test_every_message_type_round_trips(mt):
    payload = {"k": "v", "n": 3, "b": b"\x00\x01", "l": [1, 2, 3]}
    frame = P.encode_message(mt, payload)
    assert int.from_bytes(frame[:4], "big") == len(frame) - 4
    kind, body = P.decode_message(frame[4:]) if False else P.decode_message(frame)
    assert kind == mt and body == payload


@pytest.mark.parametrize("name,value", [("PING", 0), ("PONG", 1), ("SEARCH_REQUEST", 10), ("SEARCH_RESPONSE", 11), ("INDEX_PUBLISH", 20), ("INDEX_QUERY", 22),
                                         ("CRAWL_ASSIGN", 30), ("CRAWL_LOCK", 32), ("CRAWL_UNLOCK", 34), ("REPLICATE_REQUEST", 40), ("ATTESTATION_PUBLISH", 50),
                                         ("KEY_REVOCATION", 60), ("CREDIT_PROOF_REQUEST", 70), ("CREDIT_SYNC_ANNOUNCE", 72), ("INDEX_SUBMIT", 80),
                                         ("PEX_REQUEST", 90), ("ERROR", 99), ("SIGNED_ENVELOPE", 100)])
def test_message_type_numbers_are_wire_stable(name, value):
    assert P.MessageType[name] == value


def test_dht_keys_are_sha256_of_normalised_input():
    assert P.keyword_to_dht_key("Tensor") == "/infomesh/kw/" + hashlib.sha256(b"tensor").hexdigest()
    assert P.url_to_dht_key("https://e.org/a") == "/infomesh/url/" + hashlib.sha256(b"https://e.org/a").hexdigest()


def test_oversized_frames_are_refused():
    with pytest.raises(ValueError):
        P.read_frame_length((11 * 1024 * 1024).to_bytes(4, "big"))
    assert P.read_frame_length((1024).to_bytes(4, "big")) == 1024


# ------------------------------------------------------------------ hashing / errors / cache
@pytest.mark.parametrize("data", ["", "abc", "한국어", b"\x00\xff", "x" * 10000])
def test_content_hash_is_sha256_hex(data):
    raw = data.encode("utf-8") if isinstance(data, str) else data
    assert content_hash(data) == hashlib.sha256(raw).hexdigest() and len(short_hash(data)) == 16


@pytest.mark.parametrize("cat", list(ErrorCategory))
def test_error_catalogue_has_an_entry_per_category(cat):
    from infomesh_b200.errors import ERRORS

    mine = [(k, e) for k, e in ERRORS.items() if e.category == cat]
    assert mine, cat
    for key, e in mine:
        assert get_error(key) is e and e.code in format_error(key) and 400 <= e.http_status < 600 and e.resolution


def test_unknown_error_code_formats_gracefully():
    assert get_error("NOPE") is None and "NOPE" in format_error("NOPE")


def test_query_cache_key_ttl_and_lru():
    c = QueryCache(max_size=2, ttl_seconds=0.05)
    k1, k2, k3 = (c.make_key(q, 10) for q in ("Alpha", "beta", "gamma"))
    assert k1 == c.make_key("alpha", 10) and k1 != c.make_key("alpha", 5) and len(k1) == 16
    c.put(k1, "r1")
    c.put(k2, "r2")
    assert c.get(k1) == "r1"
    c.put(k3, "r3")                       # evicts the least recently used (k2)
    assert c.get(k2) is None and c.get(k1) == "r1" and c.stats.evictions == 1
    time.sleep(0.06)
    assert c.get(k1) is None and c.stats.misses >= 2
    assert math.isclose(c.stats.hit_rate, c.stats.hits / c.stats.total)
