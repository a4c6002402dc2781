"""config.py (precedence, coercion, validation), errors.py, slo.py, plugins.py, compression, hashing, db."""
import asyncio
import time
from pathlib import Path

import pytest

from infomesh_b200 import config as C
from infomesh_b200 import errors as E
from infomesh_b200 import plugins as PL
from infomesh_b200 import slo as SL
from infomesh_b200.compression.zstd import Compressor, ZstdError
from infomesh_b200.db import SQLiteStore
from infomesh_b200.hashing import content_hash, short_hash


# ------------------------------------------------------------------ config
def write(tmp_path, body):
    p = tmp_path / "config.toml"
    p.write_text(body, encoding="utf-8")
    return p


def test_toml_values_are_coerced_validated_and_defaulted(tmp_path, monkeypatch):
    monkeypatch.delenv("INFOMESH_CRAWL_MAX_CONCURRENT", raising=False)
    p = write(tmp_path, f'''
[node]
data_dir = "{tmp_path / "data"}"
role = "bogus"
log_level = "debug"
[crawl]
max_concurrent = 1000
politeness_delay = 2
urls_per_hour = "not a number"
[network]
replication_factor = 70
bootstrap_nodes = ["/ip4/1.2.3.4/tcp/4001/p2p/Qm"]
[storage]
compression_level = 19
[gpu]
tp = 99
''')
    cfg = C.load_config(p)
    d = C.Config()
    assert cfg.node.role == d.node.role and cfg.node.log_level == "debug"                     # not allowed -> default
    assert cfg.crawl.max_concurrent == 100                                                      # out of range -> clamped to the bound
    assert cfg.crawl.politeness_delay == 2.0 and isinstance(cfg.crawl.politeness_delay, float)  # int -> float
    assert cfg.crawl.urls_per_hour == d.crawl.urls_per_hour                                     # junk -> default
    assert cfg.network.replication_factor == 10 and cfg.network.bootstrap_nodes == ["/ip4/1.2.3.4/tcp/4001/p2p/Qm"]
    assert cfg.storage.compression_level == 19 and cfg.gpu.tp == 8
    assert cfg.index.db_path == tmp_path / "data" / "index.db" and (tmp_path / "data").is_dir()  # index follows data_dir


def test_env_overrides_beat_toml_and_parse_bools_lists(tmp_path, monkeypatch):
    p = write(tmp_path, f'[node]\ndata_dir = "{tmp_path / "d"}"\n[crawl]\nmax_concurrent = 7\n')
    monkeypatch.setenv("INFOMESH_CRAWL_MAX_CONCURRENT", "9")
    monkeypatch.setenv("INFOMESH_NETWORK_BOOTSTRAP_NODES", "/ip4/1.1.1.1/tcp/1/p2p/a, /ip4/2.2.2.2/tcp/1/p2p/b")
    monkeypatch.setenv("INFOMESH_CRAWL_RESPECT_ROBOTS", "no")
    cfg = C.load_config(p)
    assert cfg.crawl.max_concurrent == 9 and len(cfg.network.bootstrap_nodes) == 2
    if hasattr(cfg.crawl, "respect_robots"):
        assert cfg.crawl.respect_robots is False
    monkeypatch.setenv("INFOMESH_CRAWL_MAX_CONCURRENT", "abc")
    assert C.load_config(p).crawl.max_concurrent == C.CrawlConfig().max_concurrent


def test_unreadable_toml_falls_back_to_defaults(tmp_path, monkeypatch):
    monkeypatch.setenv("INFOMESH_NODE_DATA_DIR", str(tmp_path / "dd"))
    cfg = C.load_config(write(tmp_path, "this is = = not toml ["))
    assert cfg.crawl == C.CrawlConfig() and cfg.node.data_dir == tmp_path / "dd"
    assert C._coerce("~/x", Path) == Path("~/x").expanduser() and C._coerce("a, b,,c", list) == ["a", "b", "c"]
    assert C._coerce("YES", bool) is True and C._coerce("0", bool) is False and C._coerce(3, float) == 3.0


# ------------------------------------------------------------------ error catalogue
def test_error_catalogue_lookup_and_formatting():
    e = E.get_error("E005")
    assert e.category == E.ErrorCategory.CRAWL and "SSRF" in e.message and e.http_status == 400
    assert e.to_dict()["error"]["code"] == "INFOMESH_E005" and e.format().startswith("Error [INFOMESH_E005]:") and "Resolution:" in E.format_error("E001")
    assert E.get_error("E999") is None and E.format_error("E999") == "Unknown error: E999"
    assert E.get_error("E001").http_status == 401 and E.get_error("E006").http_status == 429


# ------------------------------------------------------------------ SLOs
def test_slo_ratio_and_latency_tracking():
    t = SL.SLOTracker()
    for i in range(100):
        t.record_success("search_availability", i != 0)                # 99 %
        t.record("search_latency_p99", 10.0 + i)                       # p99 = 109 ms
    for ok in (True, False, False, True):
        t.record_success("crawl_success_rate", ok)                     # 50 % < 90 %
    st = {s.slo.name: s for s in t.get_status()}
    assert st["search_availability"].met and st["search_availability"].current_value == 0.99
    assert st["search_latency_p99"].met and st["search_latency_p99"].current_value == 109.0
    assert not st["crawl_success_rate"].met and st["crawl_success_rate"].error_budget_remaining == 0.0
    assert st["node_uptime"].met and st["node_uptime"].current_value == 1.0                 # no data: assume healthy
    s = t.summary()
    assert s["total_slos"] == len(SL.DEFAULT_SLOS) and s["slos_violated"] == 1


def test_slo_latency_window_excludes_old_samples():
    slo = SL.SLODefinition("lat", "d", 50.0, "ms", window_seconds=0.05)
    t = SL.SLOTracker([slo])
    t.record("lat", 500.0)
    time.sleep(0.07)
    t.record("lat", 20.0)
    st = t.get_status()[0]
    assert st.met and st.current_value == 20.0 and st.error_budget_remaining == 0.6


# ------------------------------------------------------------------ plugins
def test_plugin_hooks_chain_filter_and_survive_errors():
    reg = PL.PluginRegistry()

    @reg.hook(PL.HookPoint.PRE_INDEX)
    def upper(doc):
        return {**doc, "title": doc["title"].upper()}

    def broken(doc):
        raise RuntimeError("boom")

    def drop_spam(doc):
        return None if "spam" in doc["title"].lower() else doc

    reg.register_plugin("filters", "1.2.0", {PL.HookPoint.PRE_INDEX: broken})
    reg.register_plugin("spam", hooks={PL.HookPoint.PRE_INDEX: drop_spam})
    assert reg.run_hook(PL.HookPoint.PRE_INDEX, {"title": "hello"}) == {"title": "HELLO"}
    assert reg.run_hook(PL.HookPoint.PRE_INDEX, {"title": "Spam offer"}) is None
    assert reg.run_hook(PL.HookPoint.POST_SEARCH, [1, 2]) == [1, 2]
    assert reg.hook_counts == {"pre_index": 3} and {p["name"] for p in reg.registered_plugins} == {"filters", "spam"}
    assert reg.unregister_plugin("spam") and not reg.unregister_plugin("spam")
    assert reg.run_hook(PL.HookPoint.PRE_INDEX, {"title": "Spam offer"}) == {"title": "SPAM OFFER"}
    assert PL.get_registry() is PL.get_registry()


def test_plugin_async_hooks():
    reg = PL.PluginRegistry()

    async def add_one(x):
        return x + 1

    reg.register_plugin("a", hooks={PL.HookPoint.PRE_RANK: add_one})
    reg.register_plugin("b", hooks={PL.HookPoint.PRE_RANK: lambda x: x * 10})
    assert asyncio.run(reg.run_hook_async(PL.HookPoint.PRE_RANK, 1)) == 20


# ------------------------------------------------------------------ compression / hashing / sqlite base
def test_compressor_levels_text_and_bomb_guard():
    c = Compressor(level=3)
    blob = ("search engines " * 500).encode()
    z = c.compress(blob)
    assert len(z) < len(blob) // 5 and c.decompress(z) == blob and c.level == 3
    assert c.decompress_text(c.compress_text("héllo wörld")) == "héllo wörld" and c.compress(b"") is not None
    with pytest.raises(ZstdError):
        c.decompress(z, max_output_size=100)
    with pytest.raises(ZstdError):
        c.decompress(b"not a zstd frame at all")


def test_hashes_and_sqlite_store(tmp_path):
    assert content_hash("abc") == content_hash(b"abc") == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    assert short_hash("abc") == content_hash("abc")[:16] and len(short_hash("abc", 8)) == 8

    class Notes(SQLiteStore):
        _SCHEMA = "CREATE TABLE IF NOT EXISTS notes (k TEXT PRIMARY KEY, v TEXT);"

    s = Notes(tmp_path / "sub" / "notes.db")
    s.conn.execute("INSERT INTO notes VALUES ('a', 'b')")
    s.conn.commit()
    assert s.conn.execute("PRAGMA journal_mode").fetchone()[0].lower() == "wal" and str(tmp_path) in s.path
    s.close()
    again = Notes(tmp_path / "sub" / "notes.db")
    assert again.conn.execute("SELECT v FROM notes WHERE k = 'a'").fetchone()[0] == "b"
    again.close()
    mem = Notes(None)
    assert mem.path in (":memory:", "") or "memory" in mem.path
