"""CLI, SDK, framework adapters, dashboard (model: reference tests/test_cli*.py, test_sdk.py, test_integrations.py,
test_dashboard*.py)."""
import asyncio
import json

import pytest
from click.testing import CliRunner


@pytest.fixture()
def node_env(tmp_path, monkeypatch):
    monkeypatch.setenv("INFOMESH_NODE_DATA_DIR", str(tmp_path))
    monkeypatch.setenv("INFOMESH_INDEX_DB_PATH", str(tmp_path / "index.db"))
    monkeypatch.setenv("INFOMESH_INDEX_VECTOR_SEARCH", "false")
    monkeypatch.setenv("INFOMESH_NETWORK_BOOTSTRAP_DNS", "false")
    monkeypatch.setenv("INFOMESH_NETWORK_BOOTSTRAP_GITHUB", "false")
    import infomesh_b200.config as C

    monkeypatch.setattr(C, "DEFAULT_CONFIG_PATH", tmp_path / "config.toml")
    return tmp_path


def _seed(data_dir):
    from infomesh_b200.sdk.client import InfoMeshClient

    with InfoMeshClient(str(data_dir)) as c:
        c.add_document("https://docs.python.org/3/library/asyncio.html", "asyncio — Asynchronous I/O",
                       "asyncio is a library to write concurrent code using the async await syntax. The event loop runs tasks. " * 3)
        c.add_document("https://doc.rust-lang.org/book/ownership.html", "Understanding Ownership",
                       "Ownership is a set of rules that govern how a Rust program manages memory. Borrowing keeps references valid. " * 3)


def test_cli_keys_without_a_key_pair_change_nothing(tmp_path, monkeypatch):
    from click.testing import CliRunner

    from infomesh_b200.cli import cli

    monkeypatch.setenv("INFOMESH_NODE_DATA_DIR", str(tmp_path / "fresh"))
    out = CliRunner().invoke(cli, ["keys", "export"])
    assert out.exit_code == 0 and "No public key found" in out.output and "Run 'infomesh start' first." in out.output
    assert "No existing key pair to rotate" in CliRunner().invoke(cli, ["keys", "rotate", "--yes"]).output
    assert not (tmp_path / "fresh" / "keys" / "public.pem").exists()          # the node mints keys on its first start, not this command


def test_cli_commands(node_env):
    from infomesh_b200.cli import cli

    _seed(node_env)
    run = CliRunner().invoke
    out = run(cli, ["search", "--local", "asyncio event loop", "-n", "3"])
    assert out.exit_code == 0 and "asyncio — Asynchronous I/O" in out.output and "docs.python.org" in out.output
    assert "No results found." in run(cli, ["search", "--local", "zzzqqq"]).output
    st = run(cli, ["index", "stats"])
    assert st.exit_code == 0 and "Documents:       2" in st.output and "Tokenizer:       unicode61" in st.output and "docs.python.org" in st.output
    snap = str(node_env / "out.infomesh-snapshot")
    exp = run(cli, ["index", "export", snap]).output
    assert f"Exported 2 documents to {snap}" in exp and "  File size: " in exp
    imp = run(cli, ["index", "import", snap]).output
    assert f"Imported 0 documents from {snap}" in imp and "  Skipped (duplicate): 2" in imp and "  Total in snapshot:   2" in imp and "  Time: " in imp
    meta = run(cli, ["index", "import", snap, "--info"]).output
    assert meta.startswith(f"Snapshot: {snap}") and "  Documents:      2" in meta and "  Format version: " in meta
    assert "Missing argument 'INPUT_PATH'" in run(cli, ["index", "import"]).output
    (node_env / "urls.txt").write_text("https://a.example/1\nhttps://a.example/2\n")
    reg = run(cli, ["index", "import-urls", str(node_env / "urls.txt")]).output
    assert "Registered 2 URLs from " in reg and "  Skipped (already seen): 0" in reg
    show = run(cli, ["config", "show"])
    assert show.exit_code == 0 and "[crawl]" in show.output and "politeness_delay" in show.output
    setr = run(cli, ["config", "set", "crawl.politeness_delay", "2.5"])
    assert setr.exit_code == 0 and "crawl.politeness_delay = 2.5" in setr.output and "politeness_delay = 2.5" in (node_env / "config.toml").read_text()
    assert run(cli, ["config", "set", "nope.key", "1"]).exit_code != 0
    assert "✔ GitHub identity set" in run(cli, ["config", "github", "dev@example.com"]).output
    assert run(cli, ["config", "github", "not-an-email"]).exit_code != 0
    from infomesh_b200.p2p.keys import ensure_keys

    ensure_keys(node_env)
    kx = run(cli, ["keys", "export"])
    assert kx.output.startswith("-----BEGIN PUBLIC KEY-----") and "END PUBLIC KEY" in kx.output        # pipeable PEM
    rot = run(cli, ["keys", "rotate", "--yes"]).output
    assert "Old keys backed up. Revocation record saved." in rot and "Old Peer ID:" in rot and "New Peer ID:" in rot
    added = run(cli, ["peer", "add", "/ip4/127.0.0.1/tcp/1"]).output
    assert "Added: /ip4/127.0.0.1/tcp/1" in added and "Warning: No /p2p/<PEER_ID>" in added and "Testing TCP 127.0.0.1:1... unreachable" in added
    assert "Already in bootstrap list." in run(cli, ["peer", "add", "/ip4/127.0.0.1/tcp/1"]).output
    pl = run(cli, ["peer", "list"])
    assert "Bootstrap nodes:\n" in pl.output and "/ip4/127.0.0.1/tcp/1" in pl.output and "P2P state: not started" in pl.output
    pt = run(cli, ["peer", "test"]).output
    assert "TCP 127.0.0.1:1 ... FAIL" in pt and "reachable" in pt.split("Result:")[1]
    assert "Removed: /ip4/127.0.0.1/tcp/1" in run(cli, ["peer", "remove", "/ip4/127.0.0.1/tcp/1"]).output
    assert "Not found in bootstrap list." in run(cli, ["peer", "remove", "/ip4/127.0.0.1/tcp/1"]).output
    assert "Invalid multiaddr format." in run(cli, ["peer", "add", "udp://x"]).output
    (node_env / "feeds.opml").write_text('<opml><body><outline text="a" xmlUrl="https://a.example/feed.xml"/></body></opml>')
    assert "1 new feeds" in run(cli, ["feeds", "import", str(node_env / "feeds.opml")]).output
    assert "https://a.example/feed.xml" in run(cli, ["feeds", "list"]).output
    assert "RSS feed monitoring is disabled." in run(cli, ["feeds", "list"]).output
    fs = run(cli, ["feedback", "stats"]).output            # the searches above may already have created feedback.db
    assert "Total signals: 0" in fs or "No feedback data yet. Search more to collect signals." in fs
    ft = run(cli, ["feedback", "top-urls"]).output
    assert "No boosted URLs yet." in ft or "No feedback data yet." in ft
    doc = run(cli, ["doctor"])
    assert doc.exit_code == 0 and "InfoMesh Doctor" in doc.output and "Summary:" in doc.output
    bench = run(cli, ["bench", "-n", "3"])
    assert bench.exit_code == 0 and "query_expansion" in bench.output and "intent_classify" in bench.output
    stat = run(cli, ["status"])
    assert stat.exit_code == 0 and "Documents:       2" in stat.output and "Running:         no" in stat.output
    assert "No running InfoMesh node found." in run(cli, ["stop"]).output
    assert run(cli, ["update", "--check"]).exit_code == 0
    txt = run(cli, ["dashboard", "--text"])
    assert txt.exit_code == 0 and "Node" in txt.output and "Documents" in txt.output and "Credits" in txt.output
    assert "blocked" in run(cli, ["crawl", "http://127.0.0.1:1/"]).output
    hidden = run(cli, ["--help"]).output
    assert "_serve" not in hidden and "start" in hidden


def test_sdk_and_framework_adapters(tmp_path):
    from infomesh_b200.integrations.haystack import HaystackDocument, InfoMeshDocumentStore
    from infomesh_b200.integrations.langchain import InfoMeshRetriever
    from infomesh_b200.integrations.llamaindex import InfoMeshReader
    from infomesh_b200.sdk import InfoMeshClient

    _seed(tmp_path)
    with InfoMeshClient(str(tmp_path), {"crawl.politeness_delay": 0.2}) as c:
        hits = c.search("rust ownership borrowing", limit=3)
        assert hits and hits[0].url.endswith("ownership.html") and hits[0].to_dict()["score"] > 0
        assert asyncio.run(c.search_async("asyncio"))[0].title.startswith("asyncio")
        assert [len(r) for r in c.search_many(["asyncio", "ownership"], limit=2)] == [1, 1]
        assert c.search("asyncio", include_domains=["doc.rust-lang.org"]) == []
        assert c.fetch_page("https://docs.python.org/3/library/asyncio.html").startswith("asyncio is a library")
        assert c.get_stats()["total_documents"] == 2 and c.suggest("Under") and c.network_info().index_size == 2
        assert c.add_document("https://docs.python.org/3/library/asyncio.html", "dup", "x" * 80) is None
        bad = c.crawl("http://127.0.0.1:1/")
        assert not bad.success and bad.error
        lc = InfoMeshRetriever(client=c, limit=2)
        docs = lc.invoke("asyncio event loop")
        assert docs[0].metadata["source"] == "infomesh" and docs[0].metadata["url"].endswith("asyncio.html") and lc.get_relevant_documents("asyncio")
        assert asyncio.run(lc.ainvoke("ownership"))[0].page_content and len(lc.batch(["asyncio", "ownership"])) == 2
        li = InfoMeshReader(client=c).load_data("ownership rules")
        assert li[0].id_.endswith("ownership.html") and next(InfoMeshReader(client=c).lazy_load_data("asyncio")).text
        hs = InfoMeshDocumentStore(client=c)
        assert hs.count_documents() == 2 and hs.query("asyncio", top_k=1)[0].score > 0
        n = hs.write_documents([HaystackDocument("Kademlia keeps k-buckets of contacts sorted by last seen time. " * 3,
                                                 {"url": "https://p2p.example/kad", "title": "Kademlia"}), HaystackDocument("", {})])
        assert n == 1 and hs.count_documents() == 3


def test_dashboard_helpers_and_tui(node_env):
    from infomesh_b200.config import load_config
    from infomesh_b200.dashboard import utils as U
    from infomesh_b200.dashboard.bgm import BGMPlayer, _build_volume_args, kill_orphaned_bgm
    from infomesh_b200.dashboard.data_cache import DashboardDataCache
    from infomesh_b200.dashboard.text_report import _make_bar, render_text_report
    from infomesh_b200.dashboard.widgets.bar_chart import render_bars
    from infomesh_b200.dashboard.widgets.resource_bar import render_resource
    from infomesh_b200.dashboard.widgets.sparkline import render_sparkline

    _seed(node_env)
    cfg = load_config()
    assert U.format_uptime(3725) == "1h 2m" and U.format_uptime(0) == "—" and U.format_uptime(90061).startswith("1d 1h") and U.format_bytes(1536) == "1.5 KB"
    assert U.read_p2p_status(cfg) == {} and not U.is_node_running(cfg) and "Tier 2" in U.tier_label(type("T", (), {"name": "TIER_2"})())
    (node_env / "p2p_status.json").write_text(json.dumps({"state": "running", "peers": 3, "timestamp": 1.0}))
    assert U.read_p2p_status(cfg)["state"] == "stopped"
    cache = DashboardDataCache(cfg, ttl=60)
    st = cache.get_stats()
    assert st.document_count == 2 and st.domain_count == 2 and st.pages_last_hour == 2 and len(st.recent_docs) == 2 and cache.get_stats() is st

    class Log:
        lines: list = []

        def log_crawl(self, line, *, success=True, credits=0):
            self.lines.append((line, success))

    log = Log()
    seen, last = U.push_new_docs_to_log(st.recent_docs, st.document_count, set(), -1, log)          # the reference's calling convention
    assert len(seen) == 2 and last == 2 and len(log.lines) == 2 and all(ok for _, ok in log.lines)
    assert log.lines[0][0].startswith("https://") and U.push_new_docs_to_log(st.recent_docs, 2, seen, last, log) == (seen, 2) and len(log.lines) == 2
    assert U.push_new_docs_to_log([], 2, seen, 2, log) == (seen, 2)
    cache.close()
    assert "95%" in _make_bar(0.95).plain and render_sparkline([1, 2, 3, 4], 4) == "▁▃▆█" and "█" in render_bars([("a.com", 3), ("b.com", 1)])
    assert "red" in render_resource("CPU", 0.95) and "no data" in render_bars([])
    rep = render_text_report(cfg)
    assert "Documents" in rep and "2" in rep and "Not started" not in rep.split("Network")[0]
    assert _build_volume_args("/usr/bin/mpv", 40) == ["--volume=40"] and _build_volume_args("afplay", 50) == ["-v", "0.50"]
    p = BGMPlayer()
    assert p.play("/nonexistent.mp3") is False and not p.is_playing and kill_orphaned_bgm() >= 0

    from infomesh_b200.dashboard.app import DashboardApp
    from infomesh_b200.dashboard.screens.search import SearchPane
    from infomesh_b200.dashboard.screens.settings import SettingsPane

    async def drive():
        app = DashboardApp(cfg, initial_tab="crawl")
        async with app.run_test(size=(120, 40)) as pilot:
            await pilot.pause()
            assert app.query_one("#tabs").active == "crawl"
            await pilot.press("3")
            assert app.query_one("#tabs").active == "search"
            text = app.query_one(SearchPane).run_query("asyncio event loop")
            assert "asyncio" in text and "docs.python.org" in text
            msg = app.query_one(SettingsPane).apply_edit("crawl.max_concurrent = 7")
            await pilot.pause()
            assert "saved" in msg and app.config.crawl.max_concurrent == 7
            assert "expected" in app.query_one(SettingsPane).apply_edit("garbage")
            await pilot.press("r")
            await pilot.press("q")
            await pilot.pause()
        return app.exit_action

    assert asyncio.run(drive()) == "dashboard_only"
