"""Crawler plane with injected fakes (no network): dedup, scheduler, robots, parser, worker pipeline, feeds, recrawl."""
from __future__ import annotations

import asyncio
import time

import pytest

from infomesh_b200.config import CrawlConfig
from infomesh_b200.crawler import content_extract, diff, feed_monitor, freshness, intelligence, js_detect, recrawl
from infomesh_b200.crawler import rss, structured
from infomesh_b200.crawler.dedup import DeduplicatorDB, normalize_url
from infomesh_b200.crawler.parser import extract_canonical, extract_content, extract_links
from infomesh_b200.crawler.robots import RobotsChecker
from infomesh_b200.crawler.scheduler import Scheduler
from infomesh_b200.crawler.simhash import SimHashIndex, hamming_distance, simhash, simhash_many
from infomesh_b200.crawler.url_assigner import UrlAssigner
from infomesh_b200.crawler.worker import CrawlWorker
from infomesh_b200.security import SSRFError, validate_url

PAGE = """<html lang="en"><head><title>GPU Search</title>
<link rel="alternate" type="application/rss+xml" href="/feed.xml"></head><body><article>
<p>Blackwell tensor cores accelerate the dense retrieval stage of a search engine by a wide margin.</p>
<p>The inverted index lives in HBM and posting lists are intersected by a CUDA kernel. <a href="/docs/next.html">next</a>
<a href="https://other.example.org/page">other</a> <a href="/file.pdf">pdf</a></p></article></body></html>"""


class Resp:
    def __init__(self, text="", status=200, headers=None, url=""):
        self.text, self.status_code, self.url = text, status, url
        self.headers = {"content-type": "text/html; charset=utf-8", **(headers or {})}


class FakeClient:
    is_closed = False

    def __init__(self, routes):
        self.routes, self.calls = routes, []

    async def get(self, url, **kw):
        self.calls.append(url)
        r = self.routes.get(url)
        if r is None:
            return Resp("", 404, url=url)
        if isinstance(r, Exception):
            raise r
        if isinstance(r, list):
            r = r.pop(0) if len(r) > 1 else r[0]
        r.url = r.url or url
        return r

    async def aclose(self):
        pass


def make_worker(routes, **cfg):
    sched = Scheduler(politeness_delay=0.0, urls_per_hour=0, pending_per_domain=100)
    dedup = DeduplicatorDB()
    client = FakeClient(routes)
    w = CrawlWorker(CrawlConfig(**cfg), sched, dedup, RobotsChecker("InfoMesh"), client_factory=lambda: client,
                    resolve_dns=False)
    return w, sched, dedup, client


def test_ssrf_guard():
    assert validate_url("https://example.com/x") == "https://example.com/x"
    for bad in ("ftp://example.com", "http://localhost/", "http://127.0.0.1/", "http://10.1.2.3/", "http://[::1]/",
                "http://169.254.169.254/latest", "http://foo.internal/", "http://192.168.1.1/", "", "http:///x",
                "http://[::ffff:127.0.0.1]/", "https://a.com/" + "x" * 3000):
        with pytest.raises(SSRFError):
            validate_url(bad)


def test_normalize_url_and_dedup(tmp_path):
    assert normalize_url("HTTPS://Example.COM/a/?utm_source=x&b=2&a=1#frag") == "https://example.com/a?a=1&b=2"
    assert normalize_url("http://x.org") == "http://x.org/"
    db = DeduplicatorDB(str(tmp_path / "d.db"))
    text = "the quick brown fox jumps over the lazy dog and keeps running through the forest " * 5
    assert not db.is_url_seen("https://a.com/p") and not db.is_near_duplicate(text)
    db.mark_seen("https://a.com/p?utm_medium=z", "h1", text)
    assert db.is_url_seen("https://A.com/p/") and db.is_content_seen("h1")
    assert db.is_near_duplicate(text + " extra") and not db.is_near_duplicate("completely different words here " * 9)
    db.close()
    again = DeduplicatorDB(str(tmp_path / "d.db"))                  # SimHash index survives a restart
    assert again.simhash_index.size == 1 and again.is_near_duplicate(text)


def test_simhash_api_and_index():
    base = " ".join(f"word{i}" for i in range(400))
    a, b = simhash(base), simhash(base + " tail")
    assert hamming_distance(a, b) <= 10 and hamming_distance(a, simhash("totally unrelated content " * 50)) > 10 and simhash("") == 0
    assert simhash_many(["one two three", ""], device=None) == [simhash("one two three"), 0]
    idx = SimHashIndex(max_entries=3)
    for i, fp in enumerate([0b1111, 0b1011, 0xFFFF0000, 0xABCDEF]):
        idx.add(i, fp)
    assert idx.size == 3 and idx.find_near_duplicates(0b1011, threshold=1) == [1]     # first entry evicted (FIFO)
    assert idx.find_near_duplicates_batch([0xFFFF0001, 0x1], threshold=1) == [True, False]
    idx.remove(1, 0b1011)
    assert idx.get_stats() == {"unique_fingerprints": 2, "total_documents": 2}


def test_scheduler_limits():
    async def go():
        s = Scheduler(politeness_delay=0.05, urls_per_hour=0, pending_per_domain=2, max_depth=2)
        assert await s.add_url("https://a.com/1") and await s.add_url("https://a.com/2")
        assert not await s.add_url("https://a.com/3")              # per-domain cap
        assert not await s.add_url("https://b.com/deep", depth=3)  # depth cap
        t0 = time.monotonic()
        u1, _ = await s.get_url()
        u2, _ = await s.get_url()
        assert time.monotonic() - t0 >= 0.04 and (u1, u2) == ("https://a.com/1", "https://a.com/2")
        s.mark_done(u1); s.mark_error(u2)
        assert s.domain_state("a.com").pending_count == 0 and s.domain_state("a.com").error_count == 1
        s.set_crawl_delay("a.com", 500)
        assert s.domain_state("a.com").crawl_delay == 60.0
    asyncio.run(go())


def test_robots_checker():
    async def go():
        rc = RobotsChecker("InfoMesh")
        client = FakeClient({"https://a.com/robots.txt": Resp("User-agent: *\nDisallow: /private\nCrawl-delay: 2.5\n"
                                                              "Sitemap: https://a.com/sitemap.xml\n"),
                             "https://err.com/robots.txt": OSError("boom")})
        assert await rc.is_allowed(client, "https://a.com/public") and not await rc.is_allowed(client, "https://a.com/private/x")
        assert client.calls.count("https://a.com/robots.txt") == 1          # cached
        assert rc.get_crawl_delay("a.com") == 2.5 and rc.get_sitemaps("a.com") == ["https://a.com/sitemap.xml"]
        assert await rc.is_allowed(client, "https://err.com/x") and await rc.is_allowed(client, "https://none.com/x")
    asyncio.run(go())


def test_parser_extracts_text_links_canonical():
    p = extract_content(PAGE, "https://ex.com/docs/a.html")
    assert p.title == "GPU Search" and p.language == "en" and "posting lists" in p.text and "next" in p.text
    assert extract_content("<html><body><p>tiny</p></body></html>", "https://ex.com") is None
    assert extract_links(PAGE, "https://ex.com/docs/a.html") == ["https://ex.com/docs/next.html",
                                                                  "https://other.example.org/page"]
    assert extract_canonical('<link href="https://ex.com/c" rel="canonical">', "https://ex.com/x") == "https://ex.com/c"
    assert extract_canonical("<p>none</p>", "https://ex.com/x") is None


def test_worker_pipeline_success_and_rejections():
    async def go():
        routes = {"https://ex.com/docs/a.html": Resp(PAGE),
                  "https://ex.com/robots.txt": Resp("User-agent: *\nDisallow: /secret\n"),
                  "https://ex.com/docs/dup.html": Resp(PAGE.replace("GPU Search", "Other title")),
                  "https://ex.com/img": Resp("x", headers={"content-type": "image/png"}),
                  "https://ex.com/flaky": [Resp("", 503), Resp("", 503), Resp("", 503)],
                  "https://ex.com/canon": Resp(PAGE.replace("</head>", '<link rel="canonical" href="https://ex.com/real"></head>')
                                               .replace("Blackwell", "Hopper and Blackwell"))}
        w, sched, dedup, client = make_worker(routes)
        w.set_scope("https://ex.com/docs/")
        r = await w.crawl_url("https://ex.com/docs/a.html")
        assert r.success and r.page.title == "GPU Search" and r.discovered_feeds == ["https://ex.com/feed.xml"]
        assert sched.pending_count == 1                                  # only the in-scope link was scheduled
        assert (await w.crawl_url("https://ex.com/docs/a.html")).error == "already_seen"
        assert (await w.crawl_url("https://ex.com/docs/dup.html")).error == "duplicate_content"
        assert (await w.crawl_url("https://ex.com/secret/x")).error == "blocked_by_robots"
        assert (await w.crawl_url("http://127.0.0.1/x")).error.startswith("blocked:")
        assert (await w.crawl_url("https://ex.com/img")).error.startswith("unsupported_content_type")
        assert (await w.crawl_url("https://ex.com/missing")).error == "http_404"
        assert (await w.crawl_url("https://ex.com/canon")).error == "canonical_redirect:https://ex.com/real"
        assert (await w.crawl_url("https://ex.com/docs/a.html", force=True)).success
    import infomesh_b200.crawler.worker as W
    W._RETRY_BACKOFF_BASE = 0.0
    asyncio.run(go())


def test_worker_retries_5xx_then_gives_up_and_dht_lock():
    class Lock:
        def __init__(self, ok):
            self.ok, self.released = ok, 0

        async def acquire_crawl_lock(self, url):
            return self.ok

        async def release_crawl_lock(self, url):
            self.released += 1

    async def go():
        import infomesh_b200.crawler.worker as W
        W._RETRY_BACKOFF_BASE = 0.0
        routes = {"https://ex.com/flaky": [Resp("", 503), Resp("", 502), Resp(PAGE)]}
        w, *_ , client = make_worker(routes, respect_robots=False)
        assert (await w.crawl_url("https://ex.com/flaky")).success and client.calls.count("https://ex.com/flaky") == 3
        w2, *_ = make_worker({}, respect_robots=False)
        w2._dht = Lock(False)
        assert (await w2.crawl_url("https://ex.com/x")).error == "locked_by_peer"
        w3, *_ = make_worker({"https://ex.com/a": Resp(PAGE)}, respect_robots=False)
        w3._dht = Lock(True)
        assert (await w3.crawl_url("https://ex.com/a")).success and w3._dht.released == 1
    asyncio.run(go())


def test_js_detect_and_structured_and_tables():
    spa = '<html><body><div id="root"></div><noscript>Please enable JavaScript</noscript><script>window.__NEXT_DATA__={}</script></body></html>'
    d = js_detect.detect_js_requirement(spa)
    assert d.js_required and d.confidence >= 0.9 and len(d.signals) >= 4
    assert not js_detect.detect_js_requirement(PAGE).js_required
    sd = structured.extract_structured_data('<script type="application/ld+json">{"@type":"Article"}</script>'
                                            '<meta property="og:title" content="T"><meta name="keywords" content="a, b">'
                                            '<meta name="description" content=" D ">')
    assert sd.json_ld == [{"@type": "Article"}] and sd.opengraph == {"title": "T"} and sd.meta_keywords == ["a", "b"]
    assert sd.meta_description == "D" and bool(sd)
    t = content_extract.extract_tables("<table><caption>C</caption><tr><th>k</th><th>v</th></tr><tr><td>a</td><td>1 &amp; 2</td></tr></table>")[0]
    assert t.headers == ["k", "v"] and t.to_dict_list() == [{"k": "a", "v": "1 & 2"}] and t.caption == "C"
    cb = content_extract.extract_code_blocks('<pre><code class="language-Python">a = 1\nb = &lt;2&gt;</code></pre>')[0]
    assert cb.language == "python" and cb.line_count == 2 and "<2>" in cb.code


def test_feeds_opml_monitor_and_freshness_queue():
    fm = feed_monitor.FeedMonitor()
    n = fm.add_feeds_from_opml('<opml><body><outline text="Blog" xmlUrl="https://b.org/rss"/><outline xmlUrl="https://b.org/rss"/>'
                               '<outline title="News" xmlUrl="https://n.org/atom" /></body></opml>')
    assert n == 2 and fm.feeds[0].label == "Blog"
    fm.add_feed("https://sec.org/feed", priority=feed_monitor.FeedPriority.CRITICAL)
    assert fm.get_due_feeds()[0].url == "https://sec.org/feed"
    xml = "<rss><channel><title>T</title><item><title>A</title><link>https://b.org/a</link></item><item><link>https://b.org/b</link></item></channel></rss>"
    up = fm.process_feed_response("https://b.org/rss", xml, now=1000.0)
    assert up.new_urls == ["https://b.org/a", "https://b.org/b"]
    assert fm.process_feed_response("https://b.org/rss", xml, now=1001.0).new_urls == []
    assert [f.url for f in fm.get_due_feeds(now=1000.0 + 100)] == ["https://sec.org/feed", "https://n.org/atom"]
    assert fm.process_feed_response("https://unknown", xml).error == "feed not registered"
    assert fm.stats.total_new_urls == 2 and fm.stats.feeds_by_priority["critical"] == 1
    assert rss.discover_feeds(PAGE, "https://ex.com/x") == ["https://ex.com/feed.xml"]
    q = freshness.PriorityRecrawlQueue(max_size=3)
    assert q.enqueue("u1", freshness.RecrawlTrigger.SCHEDULED, now=1) and q.enqueue("u2", freshness.RecrawlTrigger.USER_REQUEST, now=2)
    assert not q.enqueue("u1", freshness.RecrawlTrigger.RSS_UPDATE) and q.enqueue("u3", freshness.RecrawlTrigger.RSS_UPDATE, now=3)
    assert not q.enqueue("u4", freshness.RecrawlTrigger.RSS_UPDATE)      # full
    q.discard("u3")
    assert [q.dequeue().url, q.dequeue().url, q.dequeue()] == ["u2", "u1", None]
    assert freshness.classify_freshness(100, now=100 + 7200) == "warm"
    assert freshness.ConditionalHeaders('"e"', "Mon").to_request_headers() == {"If-None-Match": '"e"', "If-Modified-Since": "Mon"}
    assert freshness.ConditionalHeaders.from_response_headers({"ETag": "x", "Last-Modified": "y"}).last_modified == "y"


def test_recrawl_outcomes_and_intervals():
    assert [recrawl.compute_recrawl_interval(f) for f in (0.0, 0.05, 0.3, 0.9)] == [2592000, 604800, 86400, 21600]
    assert recrawl.update_change_frequency(0.5, True) == pytest.approx(0.65)

    async def go():
        from infomesh_b200.hashing import content_hash
        c = FakeClient({"https://a.com/304": Resp("", 304), "https://a.com/same": Resp("body"),
                        "https://a.com/new": Resp("new body", headers={"etag": "v2"}), "https://a.com/gone": Resp("", 410)})
        o = await recrawl.recrawl_url("https://a.com/304", "e", None, "h", 2, client=c)
        assert (o.status, o.stale_count) == ("not_modified", 0)
        assert (await recrawl.recrawl_url("https://a.com/same", None, None, content_hash("body"), 0, client=c)).status == "not_modified"
        o = await recrawl.recrawl_url("https://a.com/new", None, None, "old", 0, client=c)
        assert o.status == "updated" and o.new_etag == "v2" and o.new_text == "new body"
        assert (await recrawl.recrawl_url("https://a.com/gone", None, None, "h", 2, client=c)).status == "deleted"
        assert (await recrawl.recrawl_url("https://a.com/gone", None, None, "h", 0, client=c)).status == "error"
        assert (await recrawl.recrawl_url("http://10.0.0.1/", None, None, "h", 0, client=c)).status == "error"
    asyncio.run(go())
    now = 1_000_000.0
    docs = [recrawl.RecrawlCandidate(i, f"u{i}", "h", None, None, 3600, sc, 0.0, now - age, None)
            for i, (age, sc) in enumerate([(7200, 0), (100, 0), (99999, 3), (4000, 0)])]
    assert [d.url for d in recrawl.select_candidates(docs, now=now)] == ["u0", "u3"]


def test_diff_warc_intelligence_assigner():
    d = diff.compute_diff("a\nb\nc", "a\nc\nd", "u")
    assert d.has_changed and d.added_lines == ["d"] and d.removed_lines == ["b"] and d.change_ratio == 0.5
    assert not diff.compute_diff("x", "x").has_changed
    warc = diff.export_warc_file([{"url": "https://a", "text": "héllo", "crawled_at": 0}, {"url": "", "text": "x"}])
    assert warc.count("WARC/1.0") == 2 and "Content-Length: 6" in warc and "WARC-Target-URI: https://a" in warc
    rc = intelligence.RobotsCache(ttl_seconds=100)
    rc.put("a.com", True, 1.5, ["s"])
    other = intelligence.RobotsCache()
    assert other.import_from_dht(rc.export_for_dht()) == 1 and other.get("a.com").crawl_delay == 1.5
    tuner = intelligence.CrawlSpeedTuner(base_delay=1.0)
    assert tuner.adjust(cpu=95, mem=10).current_delay == 1.5 and tuner.adjust(cpu=10, mem=10).current_delay == 1.2
    assert intelligence.extract_image_alt_texts('<img alt="logo"><img src=x alt="Diagram of a GPU">') == ["Diagram of a GPU"]
    ua = UrlAssigner("me")
    for p in ("p1", "p2", "p3"):
        ua.add_peer(p)
    owners = {ua.closest_peer(f"https://x/{i}") for i in range(50)}
    assert len(owners) >= 3 and ua.known_peers == 4
    assert ua.filter_local_urls([f"https://x/{i}" for i in range(50)]) == [u for u in (f"https://x/{i}" for i in range(50)) if ua.is_local_owner(u)]
