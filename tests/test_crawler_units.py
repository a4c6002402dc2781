"""crawler/robots.py, url_assigner.py, diff.py, js_detect.py, structured.py, rss.py, dedup.py, parser.py — pure units."""
import asyncio

import pytest

from infomesh_b200.crawler import diff as DF
from infomesh_b200.crawler import rss as RS
from infomesh_b200.crawler.dedup import DeduplicatorDB, normalize_url
from infomesh_b200.crawler.js_detect import detect_js_requirement
from infomesh_b200.crawler.parser import extract_canonical, extract_content, extract_links, extract_main_text
from infomesh_b200.crawler.robots import RobotsChecker
from infomesh_b200.crawler.structured import extract_structured_data
from infomesh_b200.crawler.url_assigner import UrlAssigner

ROBOTS = """User-agent: *
Disallow: /private/
Crawl-delay: 2.5
Sitemap: https://ex.org/sitemap.xml
Sitemap: https://ex.org/news.xml

User-agent: BadBot
Disallow: /
"""


# ------------------------------------------------------------------ robots
def test_parse_robots_rules_sitemaps_delay():
    parser, sitemaps, delay = RobotsChecker.parse_robots(ROBOTS, "https://ex.org/robots.txt")
    assert parser.can_fetch("InfoMesh", "https://ex.org/public") and not parser.can_fetch("InfoMesh", "https://ex.org/private/x")
    assert not parser.can_fetch("BadBot", "https://ex.org/public")
    assert sitemaps == ["https://ex.org/sitemap.xml", "https://ex.org/news.xml"] and delay == 2.5
    assert RobotsChecker.parse_robots("")[2] is None


class FakeClient:
    def __init__(self, status=200, text=ROBOTS, boom=False):
        self.status, self.text, self.boom, self.calls = status, text, boom, 0

    async def get(self, url, **kw):
        self.calls += 1
        if self.boom:
            raise OSError("tls failure")
        from types import SimpleNamespace
        return SimpleNamespace(status_code=self.status, text=self.text)


def test_robots_checker_caches_per_domain_and_fails_open():
    rc = RobotsChecker("InfoMesh")
    c = FakeClient()
    assert asyncio.run(rc.is_allowed(c, "https://ex.org/a")) and not asyncio.run(rc.is_allowed(c, "https://ex.org/private/b"))
    assert c.calls == 1 and rc.get_crawl_delay("ex.org") == 2.5 and len(rc.get_sitemaps("ex.org")) == 2
    for client in (FakeClient(status=404), FakeClient(boom=True)):
        assert asyncio.run(RobotsChecker("InfoMesh").is_allowed(client, "https://other.org/private/x"))
    rc.prime("primed.org", "User-agent: *\nDisallow: /")
    assert not asyncio.run(rc.is_allowed(FakeClient(boom=True), "https://primed.org/x"))
    rc.clear_cache()
    assert rc.get_sitemaps("ex.org") == []


# ------------------------------------------------------------------ url ownership
def test_url_assigner_is_deterministic_and_partitions_urls():
    a, b = UrlAssigner("peerA"), UrlAssigner("peerB")
    for x in (a, b):
        x.add_peer("peerA"), x.add_peer("peerB"), x.add_peer("peerC")
    urls = [f"https://ex.org/{i}" for i in range(200)]
    assert all(a.closest_peer(u) == b.closest_peer(u) for u in urls)
    mine_a, mine_b = set(a.filter_local_urls(urls)), set(b.filter_local_urls(urls))
    assert mine_a and mine_b and not (mine_a & mine_b) and len(mine_a | mine_b) < 200       # peerC owns the rest
    a.remove_peer("peerC"), a.remove_peer("peerA")
    assert a.known_peers == 2 and a.assign("https://ex.org/1", depth=2).assigner_peer_id == "peerA"


# ------------------------------------------------------------------ diff / WARC
def test_compute_diff_and_warc_export():
    assert not DF.compute_diff("same", "same").has_changed
    d = DF.compute_diff("a\nb\nc", "a\nc\nd\n\n", url="u")
    assert d.has_changed and d.added_lines == ["d"] and d.removed_lines == ["b"] and d.change_ratio == 0.4
    rec = DF.export_warc_record("https://ex.org/é", "héllo", 0.0)
    assert "WARC-Target-URI: https://ex.org/é" in rec and "Content-Length: 6" in rec and "WARC-Date: 1970-01-01T00:00:00Z" in rec
    warc = DF.export_warc_file([{"url": "https://a", "text": "x", "crawled_at": 1.0}, {"url": "https://b", "text": "y", "crawled_at": 2.0}])
    assert warc.count("WARC/1.0") == 3 and "warcinfo" in warc


# ------------------------------------------------------------------ JS detection
def test_js_detection_spa_vs_article():
    spa = '<html><head><script>window.__NEXT_DATA__={}</script></head><body><div id="__next"></div><noscript>You need to enable JavaScript to run this app.</noscript></body></html>'
    r = detect_js_requirement(spa)
    assert r.js_required and r.confidence >= 0.75 and len(r.signals) >= 3
    article = "<html><body><article>" + "Plain readable text about search engines. " * 40 + "</article></body></html>"
    r2 = detect_js_requirement(article)
    assert not r2.js_required and r2.confidence == 0.0 and r2.signals == []


# ------------------------------------------------------------------ structured data
def test_structured_data_extraction():
    html = '''<head><script type="application/ld+json">{"@type": "Article", "headline": "H"}</script>
    <script type="application/ld+json">[{"@type": "Person"}, 5]</script><script type="application/ld+json">{broken</script>
    <meta property="og:title" content="OG Title"><meta property="og:type" content="article">
    <meta name="description" content=" A description. "><meta name="keywords" content="a, b ,,c"></head>'''
    sd = extract_structured_data(html)
    assert [d["@type"] for d in sd.json_ld] == ["Article", "Person"] and sd.opengraph == {"title": "OG Title", "type": "article"}      # keys without the og: prefix
    assert sd.meta_description == "A description." and sd.meta_keywords == ["a", "b", "c"] and bool(sd)
    assert not extract_structured_data("<p>nothing</p>") and sd.to_dict()["meta_keywords"] == ["a", "b", "c"]


# ------------------------------------------------------------------ feeds
ATOM = '''<feed xmlns="http://www.w3.org/2005/Atom"><title>Atom &amp; Co</title>
<entry><title>E1</title><link rel="alternate" href="/posts/1"/><summary><![CDATA[<b>bold</b> text]]></summary>
<updated>2030-01-01T00:00:00Z</updated><author><name>Ann</name></author></entry>
<entry><title>no link</title></entry></feed>'''


def test_parse_atom_and_rss_and_discovery():
    f = RS.parse_feed_xml(ATOM, "https://ex.org/feed")
    assert f.feed_type == "atom" and f.title == "Atom & Co" and len(f.items) == 1
    it = f.items[0]
    assert (it.url, it.summary, it.author, it.published) == ("https://ex.org/posts/1", "bold text", "Ann", "2030-01-01T00:00:00Z")
    rss = '<rss><channel><title>R</title><item><title>T</title><guid>https://ex.org/g</guid><description>d</description><dc:creator>Bob</dc:creator></item></channel></rss>'
    r = RS.parse_feed_xml(rss, "https://ex.org/rss")
    assert r.feed_type == "rss" and r.items[0].url == "https://ex.org/g" and r.items[0].author == "Bob"
    assert RS.parse_feed_xml("<html/>", "u").feed_type == "unknown"
    html = '<link rel="alternate" type="application/rss+xml" href="/feed.xml"><link type="application/atom+xml" href="https://o.org/a"><link type="application/rss+xml" href="/feed.xml">'
    assert RS.discover_feeds(html, "https://ex.org/blog/") == ["https://ex.org/feed.xml", "https://o.org/a"]


# ------------------------------------------------------------------ dedup
def test_normalize_url_rules():
    assert normalize_url("HTTPS://Ex.ORG/Path/?b=2&utm_source=x&a=1#frag") == "https://ex.org/Path?a=1&b=2"
    assert normalize_url("https://ex.org") == "https://ex.org/" and normalize_url("https://ex.org/") == "https://ex.org/"
    assert normalize_url("https://ex.org/a/?fbclid=zzz") == "https://ex.org/a"


def test_deduplicator_three_levels(tmp_path):
    d = DeduplicatorDB(tmp_path / "dedup.db")
    text = "the quick brown fox jumps over the lazy dog " * 20
    assert not d.is_url_seen("https://ex.org/a")
    d.mark_seen("https://ex.org/a/?utm_medium=x", "hash1", text)
    assert d.is_url_seen("https://EX.org/a") and d.is_content_seen("hash1") and not d.is_content_seen("hash2")
    assert d.is_near_duplicate(text + " extra") and not d.is_near_duplicate("completely different content about cooking pasta " * 20)
    assert d.count() == 1
    d.close()
    again = DeduplicatorDB(tmp_path / "dedup.db")
    assert again.is_url_seen("https://ex.org/a") and again.is_near_duplicate(text)
    again.close()


# ------------------------------------------------------------------ parser
PAGE = """<html><head><title> My   Page </title><link rel="canonical" href="/canon"><style>p{}</style></head><body>
<nav><a href="/nav">nav link</a></nav><script>var x = 1;</script>
<article><h1>Heading</h1><p>First paragraph with <a href="page2.html#frag">a link</a> and enough words to matter for extraction.</p>
<p>Second paragraph, also long enough to be kept by the extractor because it has many words in it.</p></article>
<a href="mailto:x@y.z">mail</a><a href="javascript:void(0)">js</a><a href="https://other.org/x?utm_source=s">ext</a><footer>footer text</footer></body></html>"""


def test_parser_text_title_links_canonical():
    text = extract_main_text(PAGE)
    assert "First paragraph" in text and "var x" not in text and "p{}" not in text
    page = extract_content(PAGE, "https://ex.org/dir/index.html")
    assert page is not None and page.title == "My Page" and "Second paragraph" in page.text
    links = extract_links(PAGE, "https://ex.org/dir/index.html")
    assert "https://ex.org/dir/page2.html" in links and all(not link.startswith(("mailto:", "javascript:")) for link in links)
    assert any(link.startswith("https://other.org/x") for link in links)
    assert extract_canonical(PAGE, "https://ex.org/dir/index.html") == "https://ex.org/canon"
    assert extract_content("<html><body><p>tiny</p></body></html>", "https://ex.org/") is None
