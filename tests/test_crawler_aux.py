"""crawler/lang_detect.py, freshness.py, intelligence.py, feed_monitor.py, recrawl.py — pure unit tests."""
import time
from unittest.mock import MagicMock

import pytest

from infomesh_b200.crawler import feed_monitor as FM
from infomesh_b200.crawler import freshness as FR
from infomesh_b200.crawler import intelligence as IN
from infomesh_b200.crawler import recrawl as RC
from infomesh_b200.crawler.lang_detect import detect_language


# ------------------------------------------------------------------ language detection
@pytest.mark.parametrize("text,lang", [
    ("The quick brown fox jumps over the lazy dog and it was not what they said", "en"),
    ("El rápido zorro marrón salta sobre el perro perezoso y no es lo que se dice para los niños", "es"),
    ("Der schnelle braune Fuchs springt über den faulen Hund und das ist nicht was sie sagen", "de"),
    ("Le renard brun rapide saute par dessus le chien paresseux et ce n'est pas ce que nous avons dit", "fr"),
    ("분산 검색 엔진은 여러 노드가 협력하여 색인을 만든다", "ko"),
    ("分散型検索エンジンはノードが協力してインデックスを作ります", "ja"),
    ("分布式搜索引擎由多个节点协作建立索引，每个节点负责一部分文档。", "zh"),
    ("Быстрая коричневая лиса прыгает через ленивую собаку и это не то что он сказал", "ru"),
    ("ภาษาไทยเป็นภาษาที่มีวรรณยุกต์", "th"),
    ("محرك بحث موزع يعمل عبر عدة عقد", "ar"),
])
def test_detect_language(text, lang):
    d = detect_language(text)
    assert d.language == lang and d.confidence > 0.3


def test_detect_language_undetermined_inputs():
    empty = detect_language("")
    assert (empty.language, empty.confidence, empty.script) == ("en", 0.0, "Unknown") and detect_language("   ").confidence == 0.0
    assert detect_language("12345 !!! 67890 12345 !!! 67890").confidence == 0.0
    assert detect_language("too short").confidence == 0.0 and detect_language("too short", min_text_length=3).script == "Latin"


def test_detect_language_scripts_do_not_bleed_into_each_other():
    assert detect_language("한국어 문장입니다. 두 번째 문장도 있습니다.").script == "Hangul"
    assert detect_language("これは日本語です。二つ目の文もあります。").script == "Kana"
    assert all(detect_language(t).confidence <= 0.95 for t in ("한국어 문장입니다. 두 번째 문장도 있습니다.", "the cat and the dog are in the house with it"))


# ------------------------------------------------------------------ freshness tiers / recrawl queue
def test_classify_freshness_boundaries():
    now = 1_000_000.0
    assert FR.classify_freshness(now - 3600, now=now) == FR.FreshnessTier.HOT
    assert FR.classify_freshness(now - 3601, now=now) == FR.FreshnessTier.WARM
    assert FR.classify_freshness(now - 86401, now=now) == FR.FreshnessTier.COLD
    assert FR.classify_freshness(now - 604801, now=now) == FR.FreshnessTier.STALE


def test_priority_recrawl_queue_orders_by_trigger_then_time():
    q = FR.PriorityRecrawlQueue()
    assert q.enqueue("sched", FR.RecrawlTrigger.SCHEDULED, now=1)
    assert q.enqueue("rss2", FR.RecrawlTrigger.RSS_UPDATE, now=3)
    assert q.enqueue("rss1", FR.RecrawlTrigger.RSS_UPDATE, now=2, source_feed="f")
    assert q.enqueue("user", FR.RecrawlTrigger.USER_REQUEST, now=9)
    assert not q.enqueue("user", FR.RecrawlTrigger.SCHEDULED)          # duplicate URL
    assert q.peek().url == "user" and q.size == 4
    assert [q.dequeue().url for _ in range(4)] == ["user", "rss1", "rss2", "sched"]
    assert q.dequeue() is None and q.total_enqueued == 4 and q.total_dequeued == 4


def test_priority_recrawl_queue_discard_capacity_clear():
    q = FR.PriorityRecrawlQueue(max_size=2)
    q.enqueue("a", FR.RecrawlTrigger.SCHEDULED)
    q.enqueue("b", FR.RecrawlTrigger.SCHEDULED)
    assert not q.enqueue("c", FR.RecrawlTrigger.USER_REQUEST)           # full
    q.discard("a")
    assert q.size == 1 and q.dequeue().url == "b"
    q.enqueue("d", FR.RecrawlTrigger.SCHEDULED)
    q.clear()
    assert q.size == 0 and q.peek() is None


def test_conditional_headers_roundtrip():
    h = FR.ConditionalHeaders.from_response_headers({"ETag": '"abc"', "Last-Modified": "Tue, 01 Jan 2030 00:00:00 GMT"})
    assert h.to_request_headers() == {"If-None-Match": '"abc"', "If-Modified-Since": "Tue, 01 Jan 2030 00:00:00 GMT"}
    assert FR.ConditionalHeaders().to_request_headers() == {}


# ------------------------------------------------------------------ robots cache / tuner / alt texts
def test_robots_cache_ttl_export_import_cleanup():
    c = IN.RobotsCache(ttl_seconds=0.05)
    c.put("a.com", True, 1.5, ["https://a.com/sitemap.xml"])
    assert c.get("a.com").crawl_delay == 1.5 and c.size == 1
    exported = c.export_for_dht()
    other = IN.RobotsCache()
    assert other.import_from_dht(exported + [{"domain": ""}, {"domain": "b.org", "crawl_delay": "oops", "sitemaps": "x"}]) == 2
    assert other.import_from_dht(exported) == 0 and other.get("b.org").crawl_delay == 0.0
    time.sleep(0.06)
    assert c.get("a.com") is None
    c.put("z.com", False)
    time.sleep(0.06)
    assert c.cleanup() == 1 and c.size == 0


def test_crawl_speed_tuner_reacts_to_load():
    t = IN.CrawlSpeedTuner(base_delay=1.0, min_delay=0.5, max_delay=2.0)
    assert t.adjust(cpu=95, mem=40).current_delay == 1.5
    assert t.adjust(cpu=95, mem=40).current_delay == 2.0                # clamped at max
    assert "moderate" in t.adjust(cpu=75, mem=40).adjustment_reason
    for _ in range(10):
        s = t.adjust(cpu=5, mem=10)
    assert s.current_delay == 0.5 and t.adjust(cpu=50, mem=60).adjustment_reason == "stable"


def test_extract_image_alt_texts_skips_placeholders_and_repeats():
    html = ('<img src=a alt="A red fox in snow"><img alt="logo" src=b><IMG ALT=\'A red fox in snow\'>'
            '<img src=c alt="Diagram of the pipeline"><img alt="ab">')
    assert IN.extract_image_alt_texts(html) == ["A red fox in snow", "Diagram of the pipeline"]


# ------------------------------------------------------------------ feed monitor
OPML = '''<opml><body><outline text="Blog A" xmlUrl="https://a.com/feed.xml"/>
<outline title="B" xmlUrl='https://b.org/rss'/><outline text="dup" xmlUrl="https://a.com/feed.xml"/><outline text="folder"/></body></opml>'''
RSS = '''<?xml version="1.0"?><rss version="2.0"><channel><title>t</title>
<item><title>one</title><link>https://a.com/1</link></item><item><title>two</title><link>https://a.com/2</link></item></channel></rss>'''


def test_parse_opml_and_bulk_add():
    feeds = FM.parse_opml(OPML)
    assert [(f.url, f.label) for f in feeds] == [("https://a.com/feed.xml", "Blog A"), ("https://b.org/rss", "B")]
    m = FM.FeedMonitor()
    assert m.add_feeds_from_opml(OPML) == 2 and m.add_feeds_from_opml(OPML) == 0


def test_feed_monitor_due_ordering_and_limits():
    m = FM.FeedMonitor(max_feeds=3)
    lo = m.add_feed("lo", priority=FM.FeedPriority.LOW)
    hi = m.add_feed("hi", priority=FM.FeedPriority.HIGH)
    cr = m.add_feed("cr", priority=FM.FeedPriority.CRITICAL, poll_interval=10)
    with pytest.raises(ValueError):
        m.add_feed("overflow")
    assert [f.url for f in m.get_due_feeds(now=100)] == ["cr", "hi", "lo"]        # never polled -> all due, by priority
    lo.last_poll_at, hi.last_poll_at, cr.last_poll_at = 100, 100, 100
    assert [f.url for f in m.get_due_feeds(now=105)] == []
    assert [f.url for f in m.get_due_feeds(now=111)] == ["cr"]
    assert [f.url for f in m.get_due_feeds(now=100 + 3600)] == ["cr", "hi", "lo"]
    assert m.add_feed("hi", priority=FM.FeedPriority.LOW, label="x") is hi and hi.priority == FM.FeedPriority.LOW
    assert m.remove_feed("hi") and not m.remove_feed("hi") and m.stats.total_feeds == 2


def test_feed_monitor_processes_responses_and_dedups_urls():
    m = FM.FeedMonitor()
    m.add_feed("https://a.com/feed.xml")
    m.mark_url_seen("https://a.com/2")
    up = m.process_feed_response("https://a.com/feed.xml", RSS, now=50)
    assert up.new_urls == ["https://a.com/1"] and up.error is None
    assert m.process_feed_response("https://a.com/feed.xml", RSS, now=60).new_urls == []
    assert m.process_feed_response("nope", RSS).error == "feed not registered"
    st = m.stats
    assert st.total_polls == 2 and st.total_new_urls == 1 and st.feeds_by_priority == {"normal": 1}


# ------------------------------------------------------------------ recrawl scheduling
def cand(i, **kw):
    base = dict(doc_id=i, url=f"u{i}", text_hash="h", etag=None, last_modified=None, recrawl_interval=100, stale_count=0,
                change_frequency=0.2, crawled_at=0.0, last_recrawl_at=None)
    base.update(kw)
    return RC.RecrawlCandidate(**base)


def test_recrawl_interval_and_frequency_update():
    assert RC.compute_recrawl_interval(0.0) == RC.INTERVAL_STATIC
    assert RC.compute_recrawl_interval(0.05) == RC.INTERVAL_LOW
    assert RC.compute_recrawl_interval(0.5) == RC.INTERVAL_MEDIUM
    assert RC.compute_recrawl_interval(0.9) == RC.INTERVAL_HIGH
    assert RC.update_change_frequency(0.5, True) == pytest.approx(0.65)
    assert RC.update_change_frequency(0.5, False) == pytest.approx(0.35)


def test_select_candidates_most_overdue_first_and_skips_stale():
    docs = [cand(1, crawled_at=0), cand(2, crawled_at=500), cand(3, last_recrawl_at=950.0),
            cand(4, crawled_at=0, stale_count=RC.STALE_THRESHOLD)]
    assert [d.doc_id for d in RC.select_candidates(docs, now=1000)] == [1, 2]
    assert [d.doc_id for d in RC.select_candidates(docs, now=1000, max_batch=1)] == [1]


def test_apply_outcome_updates_store():
    store = MagicMock()
    RC.apply_outcome(store, cand(1), RC.RecrawlOutcome("u1", "deleted"))
    store.soft_delete.assert_called_once_with("u1")
    RC.apply_outcome(store, cand(2, change_frequency=0.5), RC.RecrawlOutcome("u2", "updated", new_text_hash="n", new_text="body"), now=77)
    kw = store.update_document.call_args.kwargs
    assert kw["text"] == "body" and kw["text_hash"] == "n" and kw["last_recrawl_at"] == 77
    assert kw["change_frequency"] == pytest.approx(0.65) and kw["recrawl_interval"] == RC.INTERVAL_HIGH
    RC.apply_outcome(store, cand(3, change_frequency=0.5), RC.RecrawlOutcome("u3", "error", stale_count=2))
    kw = store.update_document.call_args.kwargs
    assert kw["text"] is None and kw["change_frequency"] == 0.5 and kw["stale_count"] == 2
