"""The continuous crawl loop, seeds and scheduler bookkeeping with fakes (model: reference tests/test_crawl_loop.py,
test_seeds.py, test_scheduler.py)."""
import asyncio
import types
from dataclasses import replace

import pytest

from infomesh_b200.config import Config
from infomesh_b200.crawler import crawl_loop as CL
from infomesh_b200.crawler.freshness import PriorityRecrawlQueue, RecrawlTrigger
from infomesh_b200.crawler.parser import ParsedPage
from infomesh_b200.crawler.scheduler import MAX_CRAWL_DELAY, Scheduler
from infomesh_b200.crawler.worker import CrawlResult
from infomesh_b200.index.local_store import LocalStore


def _page(url):
    n = abs(hash(url)) % 10_000
    return ParsedPage(url=url, title=f"T {url[-6:]}", text=f"Page {n} explains warp specialised pipelines with tensor memory, entry {url}. " * 4, language="en",
                      raw_html_hash=f"r{url}", text_hash=f"t{url}")


class Resp:
    def __init__(self, text, status=200):
        self.text, self.status_code = text, status


class Http:
    def __init__(self, pages):
        self.pages, self.gets = pages, []

    async def get(self, url, timeout=30.0):
        self.gets.append(url)
        if url not in self.pages:
            return Resp("", 404)
        return Resp(self.pages[url])


class Worker:
    def __init__(self, http=None, fail=()):
        self.http, self.fail, self.crawled = http or Http({}), set(fail), []

    async def get_http_client(self):
        return self.http

    async def crawl_url(self, url, depth=0, force=False):
        self.crawled.append((url, depth))
        if url in self.fail:
            return CrawlResult(url, False, error="http_500")
        if url.endswith("/boom"):
            raise RuntimeError("parser exploded")
        return CrawlResult(url, True, page=_page(url), discovered_links=[url + "/next"], discovered_feeds=[url + "/feed.xml"])


class Dedup:
    def __init__(self, seen=()):
        self.seen = set(seen)

    def is_url_seen(self, url):
        return url in self.seen


class Ledger:
    def __init__(self):
        self.notes = []

    def record_action(self, action, quantity, note, key_pair=None):
        self.notes.append(note)


class Gov:
    def __init__(self, pause=False, throttle=False):
        self.should_pause_crawl, self.should_throttle_crawl, self.checks = pause, throttle, 0

    def check_and_adjust(self):
        self.checks += 1
        return types.SimpleNamespace(degrade_level=types.SimpleNamespace(name="WARNING"), cpu_percent=91.0, memory_percent=40.0, throttle_factor=0.95)


def _ctx(tmp_path, **over):
    base = Config()
    cfg = replace(base, node=replace(base.node, data_dir=tmp_path), crawl=replace(base.crawl, rss_enabled=True, rss_discovery=True, rss_max_feeds=2))
    ctx = types.SimpleNamespace(config=cfg, store=LocalStore(tmp_path / "index.db"), vector_store=None, p2p_node=None, distributed_index=None,
                                worker=Worker(), scheduler=Scheduler(politeness_delay=0.0, urls_per_hour=5, pending_per_domain=100), dedup=Dedup(),
                                ledger=Ledger(), key_pair=None, governor=None, feed_monitor=None, priority_queue=None, index_submit_sender=None)
    for k, v in over.items():
        setattr(ctx, k, v)
    return ctx


# ------------------------------------------------------------------ seeds
def test_seed_files_ship_for_every_category_and_parse_comments(tmp_path):
    from infomesh_b200.crawler import seeds

    for cat in seeds.CATEGORIES:
        urls = seeds.load_seeds(cat)
        assert urls and all(u.startswith("https://") for u in urls) and len(set(urls)) == len(urls), cat
    assert len(seeds.load_seeds()) >= max(len(seeds.load_seeds(c)) for c in seeds.CATEGORIES)
    (tmp_path / "mine.txt").write_text("# header\nhttps://a.example/   # trailing note\n\nftp://nope\nhttps://a.example/\nhttp://b.example/x\n")
    assert seeds.load_seeds("mine", tmp_path) == ["https://a.example/", "http://b.example/x"]
    assert seeds.load_seeds("absent", tmp_path) == [] and seeds.load_seeds(seeds_dir=tmp_path / "nowhere") == []


# ------------------------------------------------------------------ scheduler bookkeeping
def test_scheduler_tracks_pending_errors_and_caps_crawl_delay():
    async def go():
        s = Scheduler(politeness_delay=0.0, urls_per_hour=0, pending_per_domain=2)
        assert await s.add_url("https://a.example/1") and await s.add_url("https://a.example/2") and not await s.add_url("https://a.example/3")
        assert s.pending_count == 2 and s.tracked_domains == 1 and s.domain_state("a.example").pending_count == 2
        url, depth = await s.get_url()
        s.mark_error(url)
        st = s.domain_state("a.example")
        assert (st.pending_count, st.error_count) == (1, 1) and await s.add_url("https://a.example/3")
        s.mark_done("https://never-seen.example/")                      # unknown domains are ignored
        s.set_crawl_delay("slow.example", 600)
        assert s.domain_state("slow.example").crawl_delay == MAX_CRAWL_DELAY and s.domain_state("nope") is None

    asyncio.run(go())


def test_scheduler_prunes_idle_domains_only():
    from infomesh_b200.crawler import scheduler as M

    s = Scheduler()
    s._domains["old.example"].last_request_at = -1e9
    s._domains["busy.example"].pending_count = 1
    s._domains["busy.example"].last_request_at = -1e9
    s._domains["fresh.example"].last_request_at = M.time.monotonic()
    s._prune(0)
    assert set(s._domains) == {"busy.example", "fresh.example"}


# ------------------------------------------------------------------ the loop
def test_loop_crawls_indexes_credits_and_discovers_feeds(tmp_path, monkeypatch):
    from infomesh_b200.crawler.feed_monitor import FeedMonitor

    seeds = ["https://docs.example/a", "https://docs.example/b", "https://docs.example/bad", "https://docs.example/boom", "https://docs.example/c"]
    monkeypatch.setattr(CL, "load_seeds", lambda category=None: seeds if category == "tech-docs" else [])
    ctx = _ctx(tmp_path, feed_monitor=FeedMonitor(), worker=Worker(fail={"https://docs.example/bad"}))
    n = asyncio.run(CL.seed_and_crawl_loop(ctx, "tech-docs", max_pages=3))
    assert n == 3 and ctx.store.get_stats()["document_count"] == 3
    assert [u for u, _ in ctx.worker.crawled] == seeds                  # the failed and the raising page did not stop the loop
    assert ctx.ledger.notes == ["https://docs.example/a", "https://docs.example/b", "https://docs.example/c"]
    assert len(ctx.feed_monitor.feeds) == 2                             # rss_max_feeds caps auto-discovery
    assert ctx.scheduler._per_hour == 0
    ctx.store.close()


def test_loop_is_a_noop_for_search_only_nodes(tmp_path):
    ctx = _ctx(tmp_path, worker=None)
    assert asyncio.run(CL.seed_and_crawl_loop(ctx, max_pages=1)) == 0
    ctx.store.close()


def test_seen_seeds_are_refetched_for_new_child_links(tmp_path):
    html = '<html><body><a href="/fresh">f</a><a href="/known">k</a><a href="https://other.example/x">o</a></body></html>'
    http = Http({"https://docs.example/": html})
    ctx = _ctx(tmp_path, worker=Worker(http), dedup=Dedup({"https://docs.example/", "https://docs.example/known", "https://gone.example/"}))

    async def go():
        first = await CL._enqueue_seed(ctx, "https://new.example/")
        again = await CL._enqueue_seed(ctx, "https://docs.example/")
        gone = await CL._enqueue_seed(ctx, "https://gone.example/")        # 404: nothing to rediscover
        queued = [await ctx.scheduler.get_url() for _ in range(ctx.scheduler.pending_count)]
        return first, again, gone, queued

    first, again, gone, queued = asyncio.run(go())
    assert first == (1, 0) and again == (0, 2) and gone == (0, 0)
    assert ("https://docs.example/fresh", 1) in queued and ("https://other.example/x", 1) in queued and ("https://new.example/", 0) in queued
    ctx.store.close()


def test_reseed_visits_every_category_and_needs_crawler_parts(tmp_path, monkeypatch):
    asked = []
    monkeypatch.setattr(CL, "load_seeds", lambda category=None: asked.append(category) or [f"https://{category}.example/"])
    ctx = _ctx(tmp_path)
    assert asyncio.run(CL._reseed_queue(ctx)) == len(CL.CATEGORIES) and asked == list(CL.CATEGORIES)
    ctx.scheduler = None
    assert asyncio.run(CL._reseed_queue(ctx)) == 0
    ctx.store.close()


def test_priority_queue_is_drained_in_small_batches(tmp_path):
    q = PriorityRecrawlQueue()
    for i in range(CL.PRIORITY_BATCH + 2):
        q.enqueue(f"https://news.example/{i}", RecrawlTrigger.RSS_UPDATE, source_feed="https://news.example/feed")
    q.enqueue("https://news.example/boom", RecrawlTrigger.USER_REQUEST)
    ctx = _ctx(tmp_path, priority_queue=q)
    done = asyncio.run(CL._process_priority_queue(ctx))
    assert 0 < done <= CL.PRIORITY_BATCH and q.size == CL.PRIORITY_BATCH + 3 - CL.PRIORITY_BATCH
    assert len(ctx.ledger.notes) == done and all(n.startswith("priority:") for n in ctx.ledger.notes)
    assert asyncio.run(CL._process_priority_queue(_ctx(tmp_path / "x", priority_queue=None))) == 0
    ctx.store.close()


def test_governor_backpressure_pauses_or_throttles(tmp_path, monkeypatch):
    naps = []

    async def nap(s):
        naps.append(s)

    monkeypatch.setattr(CL.asyncio, "sleep", nap)
    ctx = _ctx(tmp_path)
    assert asyncio.run(CL._apply_governor_backpressure(ctx)) is False and not naps           # no governor
    ctx.governor = Gov(pause=True)
    assert asyncio.run(CL._apply_governor_backpressure(ctx)) is True and naps == [10]
    ctx.governor = Gov(throttle=True)
    assert asyncio.run(CL._apply_governor_backpressure(ctx)) is False and naps[-1] == pytest.approx(0.1)
    ctx.store.close()


def test_dmz_crawler_submits_pages_instead_of_indexing(tmp_path):
    class Sender:
        submit_peers = ["indexer-1"]

        def __init__(self):
            self.msgs = []

        def build_submit_message(self, page, links):
            return (page.url, tuple(links))

        async def send_to_peers(self, msg):
            self.msgs.append(msg)
            return 1

    ctx = _ctx(tmp_path, index_submit_sender=Sender())
    res = CrawlResult("https://dmz.example/p", True, page=_page("https://dmz.example/p"), discovered_links=["https://dmz.example/q"])
    asyncio.run(CL._handle_crawled(ctx, res.url, res))
    assert ctx.index_submit_sender.msgs == [("https://dmz.example/p", ("https://dmz.example/q",))] and ctx.store.get_stats()["document_count"] == 0
    ctx.store.close()


def test_gpu_mirror_is_rebuilt_after_enough_new_documents(tmp_path, monkeypatch):
    class Mirror:
        _pending, rebuilds = 0, 0

        def note_added(self):
            self._pending += 1

        def rebuild(self):
            self.rebuilds += 1
            self._pending = 0

    monkeypatch.setattr(CL, "GPU_REBUILD_PENDING", 2)
    ctx = _ctx(tmp_path, gpu_index=Mirror())

    async def go():
        for i in range(5):
            await CL._index_and_publish(ctx, CrawlResult(f"https://g.example/{i}", True, page=_page(f"https://g.example/{i}")))
        await CL._index_and_publish(ctx, CrawlResult("https://g.example/0", True, page=_page("https://g.example/0")))   # duplicate: not counted

    asyncio.run(go())
    assert ctx.gpu_index.rebuilds == 2 and ctx.gpu_index._pending == 1
    ctx.store.close()
