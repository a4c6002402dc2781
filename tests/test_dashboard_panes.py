"""Dashboard panes driven headless with Textual's pilot, plus the optional-dependency seams (Playwright, PyMuPDF) and the
typing protocols (model: reference tests/test_dashboard_screens.py, test_js_render.py, test_pdf.py)."""
import asyncio
import json
import sys
import time
import types
from dataclasses import replace

import pytest

from infomesh_b200.config import Config

textual = pytest.importorskip("textual")
from textual.app import App  # noqa: E402
from textual.widgets import Static  # noqa: E402


def _cfg(tmp_path):
    base = Config()
    return replace(base, node=replace(base.node, data_dir=tmp_path), index=replace(base.index, db_path=tmp_path / "index.db", vector_search=False),
                   dashboard=replace(base.dashboard, refresh_interval=0.2))


def _seed(cfg, n=3):
    from infomesh_b200.crawler.parser import ParsedPage
    from infomesh_b200.index.local_store import LocalStore
    from infomesh_b200.services import index_document

    with LocalStore(cfg.index.db_path) as st:
        for i in range(n):
            index_document(ParsedPage(url=f"https://site{i % 2}.example/p{i}", title=f"Doc {i}", text=f"Body {i} about named barriers and mbarriers. " * 6,
                                      language="en", raw_html_hash=f"r{i}", text_hash=f"t{i}"), st)


def _host(pane_factory):
    class Host(App):
        events: list = []

        def compose(self):
            yield pane_factory()

        def on_credits_pane_credit_earned(self, msg):
            self.events.append(msg.amount)

    return Host()


def _text(widget) -> str:
    return str(widget.render())


def test_overview_and_crawl_panes_show_index_activity(tmp_path):
    from infomesh_b200.dashboard.data_cache import DashboardDataCache
    from infomesh_b200.dashboard.screens.crawl import CrawlPane
    from infomesh_b200.dashboard.screens.overview import OverviewPane
    from infomesh_b200.dashboard.widgets import LiveLog

    cfg = _cfg(tmp_path)
    _seed(cfg)
    cache = DashboardDataCache(cfg, ttl=0.0)

    async def drive():
        app = _host(lambda: OverviewPane(cfg, cache))
        async with app.run_test(size=(120, 40)) as pilot:
            await pilot.pause()
            head = _text(app.query_one("#ov-node", Static))
            assert "stopped" in head and "documents 3" in head.replace(",", "") and "domains 2" in head
            assert len(app.query_one("#ov-log", LiveLog).lines) == 3
            _seed_more(cfg)
            app.query_one(OverviewPane).refresh_data()
            await pilot.pause()
            assert len(app.query_one("#ov-log", LiveLog).lines) == 4              # only the new document is appended
        app2 = _host(lambda: CrawlPane(cfg, cache))
        async with app2.run_test(size=(120, 40)) as pilot:
            await pilot.pause()
            head = _text(app2.query_one("#cr-head", Static))
            assert "pages last hour 4" in head and "ago" in head and f"limit {cfg.crawl.urls_per_hour}/h" in head
            assert "site0.example" in _text(app2.query_one("#cr-domains"))

    def _seed_more(c):
        from infomesh_b200.crawler.parser import ParsedPage
        from infomesh_b200.index.local_store import LocalStore
        from infomesh_b200.services import index_document

        with LocalStore(c.index.db_path) as st:
            index_document(ParsedPage(url="https://site0.example/new", title="New", text="Fresh text about cluster launch control. " * 6, language="en",
                                      raw_html_hash="rn", text_hash="tn"), st)

    asyncio.run(drive())
    cache.close()


def test_network_pane_reads_the_status_file(tmp_path):
    from infomesh_b200.dashboard.screens.network import NetworkPane

    cfg = _cfg(tmp_path)

    async def drive():
        app = _host(lambda: NetworkPane(cfg))
        async with app.run_test(size=(120, 40)) as pilot:
            await pilot.pause()
            assert "not started" in _text(app.query_one("#nw-state", Static)) and "none connected" in _text(app.query_one("#nw-peers", Static))
            from infomesh_b200 import runtime as RT
            import os

            RT.write_pid_file(tmp_path, os.getpid())
            (tmp_path / "p2p_status.json").write_text(json.dumps({
                "timestamp": time.time(), "state": "running", "peers": 2, "listen_addrs": ["/ip4/0.0.0.0/tcp/4001"],
                "dht": {"keys_stored": 1234, "gets_performed": 5}, "bandwidth": {"upload_bytes": 4096, "download_bytes": 8192},
                "bootstrap": {"connected": 1, "configured": 2}, "peer_ids": ["a" * 40, "b" * 40], "peer_versions": {"a" * 40: "0.3.1"}}))
            app.query_one(NetworkPane).refresh_data()
            await pilot.pause()
            state, peers = _text(app.query_one("#nw-state", Static)), _text(app.query_one("#nw-peers", Static))
            assert "running" in state and "stored 1,234" in state and "1 connected of 2 configured" in state and "/ip4/0.0.0.0/tcp/4001" in state
            assert "v0.3.1" in peers and "v?" in peers and app.query_one(NetworkPane)._last == (4096, 8192)
            RT.clear_pid_file(tmp_path, os.getpid())

    asyncio.run(drive())


def test_credits_pane_announces_new_earnings(tmp_path):
    from infomesh_b200.credits.ledger import ActionType, CreditLedger
    from infomesh_b200.dashboard.screens.credits import CreditsPane

    cfg = _cfg(tmp_path)

    async def drive():
        app = _host(lambda: CreditsPane(cfg))
        async with app.run_test(size=(120, 40)) as pilot:
            await pilot.pause()
            assert "No credit history yet" in _text(app.query_one("#cd-head", Static))
            led = CreditLedger(tmp_path / "credits.db")
            led.record_action(ActionType.CRAWL, 4)
            app.query_one(CreditsPane).refresh_data()
            await pilot.pause()
            head = _text(app.query_one("#cd-head", Static))
            assert "balance 4.00" in head and "Tier 1" in head and "search cost" in head and app.events == []      # first sight is not an "earning"
            led.record_action(ActionType.CRAWL, 2)
            led.close()
            app.query_one(CreditsPane).refresh_data()
            await pilot.pause()
            assert app.events == [pytest.approx(2.0)] and "crawl" in _text(app.query_one("#cd-actions")).lower()

    asyncio.run(drive())


def test_live_log_is_bounded_and_escapes_markup():
    from infomesh_b200.dashboard.widgets.live_log import LiveLog

    async def drive():
        app = _host(lambda: LiveLog(max_lines=5, visible=2, id="log"))
        async with app.run_test() as pilot:
            log = app.query_one("#log", LiveLog)
            for i in range(8):
                log.write_line(f"[{i}] fetched")
            await pilot.pause()
            assert len(log.lines) == 5 and log.lines[0] == "\\[3] fetched" and _text(log).count("fetched") == 2
            log.clear_log()
            assert log.lines == []

    asyncio.run(drive())


# ------------------------------------------------------------------ optional dependencies
def test_js_renderer_degrades_without_playwright_and_drives_it_when_present(monkeypatch):
    from infomesh_b200.crawler import js_render as J

    monkeypatch.setattr(J, "is_playwright_available", lambda: False)
    r = asyncio.run(J.JSRenderer().render("https://example.org/"))
    assert not r.success and r.error == "playwright_not_installed"

    class Page:
        url = "https://example.org/final"
        closed = False

        async def goto(self, url, timeout, wait_until):
            if "slow" in url:
                raise TimeoutError("navigation timeout")

        async def content(self):
            return "<html><body>rendered</body></html>"

        async def close(self):
            Page.closed = True

    class Browser:
        pages = 0

        async def new_page(self, user_agent):
            Browser.pages += 1
            return Page()

        async def close(self):
            pass

    class PW:
        class chromium:
            @staticmethod
            async def launch(headless, args):
                assert headless and any("max-old-space-size=256" in a for a in args)
                return Browser()

        async def stop(self):
            pass

    class Starter:
        async def start(self):
            return PW()

    fake = types.ModuleType("playwright.async_api")
    fake.async_playwright = lambda: Starter()
    monkeypatch.setitem(sys.modules, "playwright", types.ModuleType("playwright"))
    monkeypatch.setitem(sys.modules, "playwright.async_api", fake)
    monkeypatch.setattr(J, "is_playwright_available", lambda: True)
    import infomesh_b200.security as sec

    monkeypatch.setattr(sec, "validate_url", lambda url, resolve_dns=False: None if "169.254" not in url else (_ for _ in ()).throw(sec.SSRFError("metadata")))

    async def go():
        jr = J.JSRenderer(max_tabs=2, max_memory_mb=256)
        ok = await jr.render("https://example.org/app")
        slow = await jr.render("https://example.org/slow")
        blocked = await jr.render("http://169.254.169.254/")
        await jr.close()
        return ok, slow, blocked, jr._browser

    ok, slow, blocked, browser = asyncio.run(go())
    assert ok.success and "rendered" in ok.html and ok.final_url.endswith("/final") and Page.closed
    assert not slow.success and "timeout" in slow.error and blocked.error.startswith("blocked") and browser is None and Browser.pages == 2


def test_pdf_extraction_with_and_without_pymupdf(monkeypatch):
    from infomesh_b200.crawler import pdf as P

    assert P.is_pdf_url("https://x.org/paper.PDF") and P.is_pdf_url("https://x.org/a.pdf/") and not P.is_pdf_url("https://x.org/pdf-guide.html")
    monkeypatch.setitem(sys.modules, "fitz", None)                       # import fitz -> ImportError
    assert P.extract_pdf_text(b"%PDF-1.7") is None

    class Doc:
        page_count = 3
        metadata = {"title": "Blackwell notes", "author": "", "pages": 3}
        closed = False

        def __getitem__(self, i):
            return types.SimpleNamespace(get_text=lambda: f"page {i}")

        def close(self):
            Doc.closed = True

    fitz = types.ModuleType("fitz")
    fitz.open = lambda stream, filetype: Doc() if stream != b"bad" else (_ for _ in ()).throw(ValueError("broken xref"))
    monkeypatch.setitem(sys.modules, "fitz", fitz)
    got = P.extract_pdf_text(b"%PDF", max_pages=2)
    assert got.text == "page 0\n\npage 1" and got.page_count == 2 and got.title == "Blackwell notes" and got.metadata == {"title": "Blackwell notes", "pages": "3"}
    assert Doc.closed and P.extract_pdf_text(b"bad") is None


def test_protocol_seams_accept_the_real_classes_and_simple_fakes():
    from infomesh_b200 import types as T
    from infomesh_b200.p2p.keys import KeyPair

    kp: T.KeyPairLike = KeyPair.generate()
    assert kp.verify(b"m", kp.sign(b"m")) and len(kp.peer_id) == 40 and len(kp.public_key_bytes()) == 32
    for name in ("peer_id", "sign", "verify", "public_key_bytes"):
        assert name in dir(T.KeyPairLike)
    for name in ("add_document", "search", "delete_document", "get_stats"):
        assert name in dir(T.VectorStoreLike)
    from infomesh_b200.index.vector_store import VectorStore

    assert all(hasattr(VectorStore, n) for n in ("add_document", "search", "delete_document", "get_stats"))


def test_module_entry_point_dispatches_to_the_cli():
    import subprocess

    out = subprocess.run([sys.executable, "-m", "infomesh_b200", "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "search" in out.stdout and "start" in out.stdout
