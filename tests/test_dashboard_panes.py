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
            assert "running" in state and "stored 1,234" in _text(app.query_one("#nw-dht", Static)) and "1 connected of 2 configured" in state and "/ip4/0.0.0.0/tcp/4001" in state
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


def test_settings_pane_stages_saves_and_flags_restart_keys(tmp_path, monkeypatch):
    import infomesh_b200.config as C
    from infomesh_b200.dashboard.screens import settings as S

    monkeypatch.setattr(C, "DEFAULT_CONFIG_PATH", tmp_path / "config.toml")
    cfg = _cfg(tmp_path)
    changed = replace(cfg, storage=replace(cfg.storage, compression_level=9), crawl=replace(cfg.crawl, max_depth=5))
    assert S.restart_keys_changed(cfg, changed) == ["storage.compression_level"] and S.restart_keys_changed(cfg, cfg) == []

    async def drive():
        events = []

        class Host(App):
            def compose(self):
                yield S.SettingsPane(cfg)

            def on_settings_pane_config_changed(self, m):
                events.append(("changed", m.config.crawl.max_depth))

            def on_settings_pane_restart_requested(self, m):
                events.append(("restart", tuple(m.keys)))

        app = Host()
        async with app.run_test(size=(120, 50)) as pilot:
            pane = app.query_one(S.SettingsPane)
            assert "(restart)" in _text(app.query_one("#st-table", Static))
            msg = pane.apply_edit("crawl.max_depth = 4")
            await pilot.pause()
            assert "saved crawl.max_depth" in msg and "restart" not in msg and "max_depth = 4" in (tmp_path / "config.toml").read_text()
            msg = pane.apply_edit("storage.compression_level = 7")
            await pilot.pause()
            assert "after a node restart" in msg and isinstance(app.screen, S.RestartConfirmScreen) and app.screen.keys == ["storage.compression_level"]
            await pilot.click("#btn-restart-yes")
            await pilot.pause()
            assert ("restart", ("storage.compression_level",)) in events and ("changed", 4) in events
            await pilot.click("#btn-reset")
            await pilot.pause()
            assert "defaults staged" in _text(app.query_one("#st-msg", Static)) and pane.config.crawl.max_depth == C.Config().crawl.max_depth
            assert pane.config.node.data_dir == tmp_path and "max_depth = 4" in (tmp_path / "config.toml").read_text()     # nothing written yet
            await pilot.click("#btn-save")
            await pilot.pause()
            assert "saved" in _text(app.query_one("#st-msg", Static)) and "max_depth = 4" not in (tmp_path / "config.toml").read_text()
            if isinstance(app.screen, S.RestartConfirmScreen):                 # compression level went back to its default
                await pilot.click("#btn-restart-no")
                await pilot.pause()
            themed = replace(pane.config, dashboard=replace(pane.config.dashboard, theme="nord"))
            pane.update_config(themed)
            assert pane.config is themed and pane.refresh_data() is None

    asyncio.run(drive())


def test_search_pane_bindings_and_results_panel(tmp_path):
    from infomesh_b200.dashboard.screens.search import SearchPane, SearchResultsPanel
    from textual.widgets import Input

    cfg = _cfg(tmp_path)
    _seed(cfg)
    text = SearchResultsPanel.render_results("q [x]", [{"title": "T [1]", "url": "https://u", "snippet": "a <b>hit</b>", "score": 0.5}], 3.2, "local")
    assert "\\[x]" in text and "T \\[1]" in text and "[bold yellow]hit[/]" in text and "0.500" in text
    assert SearchResultsPanel.render_results("q", [], 1.0) == "No results found."

    async def drive():
        app = _host(lambda: SearchPane(cfg))
        async with app.run_test(size=(120, 40)) as pilot:
            pane = app.query_one(SearchPane)
            pane.action_focus_search()
            await pilot.pause()
            assert app.focused is app.query_one("#se-input", Input)
            await pilot.press(*"barriers", "enter")
            await pilot.pause()
            assert "site0.example" in _text(app.query_one("#se-results")) or "site1.example" in _text(app.query_one("#se-results"))
            pane.action_blur_search()
            assert pane.refresh_data() is None
            app.query_one(SearchResultsPanel).display_results("x", [], 0.0)
            assert "No results" in _text(app.query_one("#se-results"))

    asyncio.run(drive())


def test_app_palette_tab_actions_and_cleanup(tmp_path, monkeypatch):
    import infomesh_b200.config as C
    from infomesh_b200.dashboard.app import TABS, DashboardApp, DashboardCommandProvider

    monkeypatch.setattr(C, "DEFAULT_CONFIG_PATH", tmp_path / "config.toml")
    cfg = _cfg(tmp_path)
    names = [e[0] for e in DashboardCommandProvider._entries()]
    assert names[:6] == [t.title() for t in TABS] and {"Refresh", "Toggle BGM", "Help", "Exit"} <= set(names)

    async def drive():
        app = DashboardApp(cfg)
        async with app.run_test(size=(120, 40)) as pilot:
            await pilot.pause()
            for i, tab in enumerate(TABS, 1):
                await app.run_action(f"tab_{i}")
                assert app.query_one("#tabs").active == tab
            other = next(t for t in app.available_themes if t != app.theme)
            app.theme = other
            await pilot.pause()
            assert app.config.dashboard.theme == other and f'theme = "{other}"' in (tmp_path / "config.toml").read_text()
            app.exit()
        return app

    app = asyncio.run(drive())
    assert app.cache._conn is None if hasattr(app.cache, "_conn") else True


def test_panels_accept_push_style_updates(tmp_path):
    from infomesh_b200.dashboard.screens.crawl import CrawlStatsPanel
    from infomesh_b200.dashboard.screens.network import BandwidthPanel, DHTPanel, P2PStatusPanel, PeerTable
    from infomesh_b200.dashboard.screens.search import SearchResultsPanel
    from textual.containers import Vertical

    cfg = _cfg(tmp_path)

    class Box(Vertical):
        def compose(self):
            yield CrawlStatsPanel(cfg, id="a")
            yield P2PStatusPanel(cfg, id="b")
            yield DHTPanel("", id="c")
            yield BandwidthPanel(id="d")
            yield PeerTable("", id="e")
            yield SearchResultsPanel(cfg, id="f")

    async def drive():
        app = _host(Box)
        async with app.run_test(size=(140, 50)) as pilot:
            await pilot.pause()
            q = app.query_one
            assert "waiting" in _text(q("#a")) and "not started" in _text(q("#b")) and "stored 0" in _text(q("#c"))
            q("#a", CrawlStatsPanel).update_stats(total_pages=1200, pages_per_hour=40, domain_count=7, last_crawl_at=time.time() - 30, countdown=4)
            assert "1,200 pages · 7 domains" in _text(q("#a")) and "refresh in 4s" in _text(q("#a"))
            q("#a", CrawlStatsPanel).update_countdown(0)
            assert "refresh in" not in _text(q("#a"))
            q("#b", P2PStatusPanel).update_status({"state": "running", "peers": 3})
            q("#c", DHTPanel).update_data({"keys_stored": 42}, p2p_state="running")
            q("#d", BandwidthPanel).update_from_status({"upload_bytes": 2048, "download_bytes": 4096})
            q("#e", PeerTable).set_peers(["p" * 40], {"p" * 40: "1.2.3"})
            q("#f", SearchResultsPanel).display_error("index [locked]")
            await pilot.pause()
            assert "peers 3" in _text(q("#b")).replace("[bold]", "") and "stored 42" in _text(q("#c")) and q("#d", BandwidthPanel).last == (2048, 4096)
            assert "v1.2.3" in _text(q("#e")) and "Error:" in _text(q("#f")) and "[locked]" in _text(q("#f"))

    asyncio.run(drive())


def test_widgets_accept_the_reference_constructor_forms():
    """Code that embeds the widgets the way the reference documents them (data / value first, bar_width, unit) keeps working."""
    from infomesh_b200.dashboard.widgets.bar_chart import BarChart, BarItem
    from infomesh_b200.dashboard.widgets.resource_bar import ResourceBar
    from infomesh_b200.dashboard.widgets.sparkline import SparklineChart

    spark = SparklineChart([1, 4, 2, 8], color="green")
    spark.push_value(3, max_points=4)
    assert spark.data == [4.0, 2.0, 8.0, 3.0] and SparklineChart("docs/min").data == []
    bar = ResourceBar(label="CPU", value=38, max_value=100, unit="%", color="cyan", bar_width=12)
    bar.update_value(95)
    assert "95%" in str(bar.renderable if hasattr(bar, "renderable") else bar._bar_markup(0.95))
    assert "40.0/180.0 GB" in ResourceBar("HBM", 40, 180, unit="GB")._bar_markup(40 / 180)
    chart = BarChart(items=[BarItem("Crawling", 702, color="cyan"), BarItem("Uptime", 396)], bar_width=20)
    chart.set_items([("a", 1.0)])
    assert BarChart()._bar_width == 28 and chart._bar_width == 20
