"""Admin API edge paths and the crawler->indexer submit bridge (model: reference tests/test_local_api.py,
tests/test_index_submit.py)."""
import asyncio
import json
import time
from dataclasses import replace
from pathlib import Path

import pytest

from infomesh_b200.config import Config
from infomesh_b200.crawler.parser import ParsedPage

fastapi = pytest.importorskip("fastapi")
from fastapi.testclient import TestClient  # noqa: E402


def _cfg(tmp_path, **net):
    base = Config()
    return replace(base, node=replace(base.node, data_dir=tmp_path), index=replace(base.index, db_path=tmp_path / "index.db", vector_search=False),
                   network=replace(base.network, **net))


def _page(i=0, text="Blackwell tensor cores keep accumulators in tensor memory. " * 5):
    from infomesh_b200.hashing import content_hash

    text = f"{text}#{i}"
    return ParsedPage(url=f"https://example.org/p{i}", title=f"Page {i}", text=text, language="en", raw_html_hash=f"raw{i}",
                      text_hash=content_hash(text))


def _store(cfg):
    from infomesh_b200.index.local_store import LocalStore

    return LocalStore(cfg.index.db_path)


# ------------------------------------------------------------------ admin API without a runtime
def test_admin_api_without_runtime_or_index(tmp_path):
    from infomesh_b200.api.local_api import create_admin_app

    cfg = _cfg(tmp_path)
    c = TestClient(create_admin_app(cfg))
    r = c.get("/readiness")
    assert r.status_code == 503 and r.json() == {"status": "not_ready", "db": "missing"}
    assert c.get("/health?detail=1").json()["db"] == "missing"
    assert c.get("/index/stats").json() == {"document_count": 0, "db_size_mb": 0.0}
    assert c.get("/index/compression").json()["avg_doc_kb"] == 0.0
    assert c.get("/credits/balance").json() == {"balance": 0.0, "total_earned": 0.0, "total_spent": 0.0}
    assert c.get("/analytics/tools").json() == {"tool_usage": {"web_search": 0, "crawl_url": 0, "fetch_page": 0, "total": 0}, "search_fetch_rate": 0.0}
    assert c.get("/docs").status_code == 404                       # interactive docs only at log level debug


def test_search_endpoint_opens_the_store_when_no_runtime_is_attached(tmp_path):
    from infomesh_b200.api.local_api import create_admin_app
    from infomesh_b200.services import index_document

    cfg = _cfg(tmp_path)
    st = _store(cfg)
    for i in range(3):
        index_document(_page(i), st)
    st.close()
    app = create_admin_app(cfg)
    c = TestClient(app)
    out = c.get("/search", params={"q": "tensor memory accumulators", "limit": 0}).json()
    assert out["total"] >= 1 and len(out["results"]) == 1          # limit clamped to [1, 20]
    assert app.state.admin.total_searches == 1 and app.state.admin.avg_latency_ms > 0
    assert c.get("/readiness").status_code == 200
    assert c.get("/index/compression").json()["documents"] == 3


def test_network_peers_reads_a_fresh_status_file_only(tmp_path):
    from infomesh_b200.api.local_api import create_admin_app

    cfg = _cfg(tmp_path)
    c = TestClient(create_admin_app(cfg))
    p = Path(tmp_path) / "p2p_status.json"
    p.write_text(json.dumps({"timestamp": time.time(), "peers": 4, "peer_id": "12D3x", "state": "running", "dht": {"keys": 9}}))
    live = c.get("/network/peers").json()
    assert live["connected"] == 4 and live["peer_id"] == "12D3x" and live["dht"] == {"keys": 9}
    p.write_text(json.dumps({"timestamp": time.time() - 120, "peers": 4}))
    assert c.get("/network/peers").json()["connected"] == 0
    p.write_text("{not json")
    assert "note" in c.get("/network/peers").json()


def test_health_and_metrics_surface_the_runtime_status_file(tmp_path):
    from infomesh_b200.api.local_api import create_admin_app
    from infomesh_b200.runtime import write_runtime_status

    cfg = _cfg(tmp_path)
    write_runtime_status(tmp_path, {"status": "running", "degrade_level": 1, "process_memory_mb": 321.5, "updated_at": time.time()})
    c = TestClient(create_admin_app(cfg))
    h = c.get("/health?detail=1").json()
    assert h["runtime"] == "running" and h["runtime_degrade_level"] == "1" and h["runtime_process_memory_mb"] == "321.5"
    assert "infomesh_process_memory_mb 321.5" in c.get("/metrics").text
    js = c.get("/metrics", headers={"accept": "application/json"}).json()
    assert "infomesh_documents_indexed" in js["metrics"]
    assert c.get("/status").json()["runtime"]["status"] == "running"


def test_config_endpoint_redacts_secrets_and_reload_failure_is_a_500(tmp_path, monkeypatch):
    from infomesh_b200.api import local_api
    from infomesh_b200.api.local_api import _redact_paths, create_admin_app

    d = {"node": {"data_dir": Path("/x"), "log_dir": Path("/var/log"), "github_email": "me@example.org"}, "llm": {"api_key": "sk-1", "token": ""},
         "network": {"bootstrap_nodes": ["/ip4/1.2.3.4/tcp/4001"]}, "n": 3}
    _redact_paths(d)
    assert d == {"node": {"data_dir": "***REDACTED***", "log_dir": "/var/log", "github_email": "***REDACTED***"},
                 "llm": {"api_key": "***REDACTED***", "token": ""}, "network": {"bootstrap_nodes": "***REDACTED***"}, "n": 3}
    bad = tmp_path / "config.toml"
    bad.write_text("[node\nbroken")
    app = create_admin_app(_cfg(tmp_path), config_path=bad)
    c = TestClient(app)
    assert c.post("/config/reload").json() == {"status": "reloaded"}      # unreadable TOML falls back to the defaults
    assert app.state.admin.config.node.data_dir != tmp_path

    def boom(path):
        raise PermissionError(path)

    monkeypatch.setattr(local_api, "load_config", boom)
    r = c.post("/config/reload")
    assert r.status_code == 500 and r.json() == {"status": "error", "detail": "Failed to reload configuration"}


def test_shell_completions_list_every_command():
    from infomesh_b200.api import extensions as X

    cmds = X.get_completion_commands()
    bash, zsh = X.generate_bash_completion(), X.generate_zsh_completion()
    from infomesh_b200.cli import cli

    assert set(cmds) == {c for c in cli.commands if not c.startswith("_")}          # completions track the real CLI
    for group, subs in X._SUBCOMMANDS.items():
        assert set(subs) == set(cli.commands[group].commands), group
    assert all(c in bash for c in cmds) and "complete -F _infomesh_complete infomesh" in bash
    assert zsh.startswith("#compdef infomesh") and "import-wet" in zsh and "gpu-build" in bash


# ------------------------------------------------------------------ index submit (sender / receiver / HTTP bridge)
def test_submit_roundtrip_with_acl_and_signature(tmp_path):
    from infomesh_b200.p2p.index_submit import IndexSubmitReceiver, IndexSubmitSender
    from infomesh_b200.p2p.keys import KeyPair
    from infomesh_b200.p2p.message_auth import PeerKeyRegistry
    from infomesh_b200.p2p.protocol import MessageType, decode_message

    crawler, stranger, indexer = KeyPair.generate(), KeyPair.generate(), KeyPair.generate()
    cfg = _cfg(tmp_path, peer_acl=[crawler.peer_id], index_submit_peers=["peer-b"])
    reg = PeerKeyRegistry()
    reg.register(crawler.peer_id, crawler.public_key_bytes())
    st = _store(cfg)
    recv = IndexSubmitReceiver(cfg, st, key_pair=indexer, key_registry=reg)
    send = IndexSubmitSender(cfg, key_pair=crawler)
    assert send.submit_peers == ["peer-b"] and recv.is_peer_allowed(crawler.peer_id) and not recv.is_peer_allowed(stranger.peer_id)

    kind, payload = decode_message(send.build_submit_message(_page(1), ["https://example.org/next"]))
    assert kind == MessageType.INDEX_SUBMIT and payload["discovered_links"] == ["https://example.org/next"] and payload["peer_id"] == crawler.peer_id
    ack = recv.handle_submit(payload)
    assert ack.success and ack.doc_id > 0 and ack.peer_id == indexer.peer_id and st.get_stats()["document_count"] == 1
    k2, a2 = decode_message(recv.build_ack_message(ack))
    assert k2 == MessageType.INDEX_SUBMIT_ACK and a2["url"] == "https://example.org/p1"

    forged = dict(payload, url="https://example.org/elsewhere")         # signature no longer covers the payload
    assert recv.handle_submit(forged).error == "bad_signature"
    swapped = dict(payload, text="attacker text under a captured signature")
    assert recv.handle_submit(swapped).error == "text_hash_mismatch"    # a sniffed signature cannot carry other text
    _, outsider = decode_message(IndexSubmitSender(cfg, key_pair=stranger).build_submit_message(_page(2)))
    assert recv.handle_submit(outsider).error == "peer_not_allowed"
    # transport-verified sender must be the claimed peer; an unsigned frame cannot claim an ACL'd identity
    assert recv.handle_submit(payload, sender=stranger.peer_id).error == "peer_id_mismatch"
    assert recv.handle_submit(payload, sender="").error == "unsigned_sender"
    assert recv.handle_submit(payload, sender=crawler.peer_id).error == ""          # duplicate url -> indexed as no-op
    reg.remove(crawler.peer_id)
    assert recv.handle_submit(payload).error == "unknown_key"          # ACL in force => a verifiable key is mandatory
    assert recv.stats["rejected"] == 6 and recv.stats["indexed"] == 2
    st.close()


def test_open_mode_accepts_unsigned_submissions_and_reports_index_errors(tmp_path):
    from infomesh_b200.p2p.index_submit import IndexSubmitReceiver, IndexSubmitSender
    from infomesh_b200.p2p.protocol import decode_message

    cfg = _cfg(tmp_path)
    st = _store(cfg)
    recv = IndexSubmitReceiver(cfg, st)
    _, payload = decode_message(IndexSubmitSender(cfg).build_submit_message(_page(3)))
    assert payload["peer_id"] == "" and payload["signature"] == b""
    assert recv.handle_submit(payload).success
    st.close()                                                         # a closed store makes indexing raise
    _, other = decode_message(IndexSubmitSender(cfg).build_submit_message(_page(4)))
    bad = recv.handle_submit(other)
    assert not bad.success and bad.error and recv.stats["indexed"] == 1


def test_sender_counts_successes_and_failures_per_peer(tmp_path):
    from infomesh_b200.p2p import index_submit as M
    from infomesh_b200.p2p.protocol import MessageType

    class Transport:
        def __init__(self):
            self.calls = []

        async def request(self, peer, kind, body, timeout):
            self.calls.append((peer, kind))
            if peer == "peer-down":
                raise ConnectionError("unreachable")
            return (MessageType.INDEX_SUBMIT_ACK, {"success": peer == "peer-ok"})

    cfg = _cfg(tmp_path, index_submit_peers=["peer-ok", "peer-refuses", "peer-down", "http://127.0.0.1:9/"])
    tr = Transport()
    s = M.IndexSubmitSender(cfg, transport=tr)
    posted = []
    orig = M._http_post
    M._http_post = lambda peer, msg: posted.append(peer) or True
    try:
        ok = asyncio.run(s.send_to_peers(s.build_submit_message(_page(4))))
    finally:
        M._http_post = orig
    assert ok == 2 and s.stats == {"sent": 2, "errors": 2} and posted == ["http://127.0.0.1:9/"]
    assert [p for p, _ in tr.calls] == ["peer-ok", "peer-refuses", "peer-down"]
    no_transport = M.IndexSubmitSender(_cfg(tmp_path, index_submit_peers=["peer-x"]))
    assert asyncio.run(no_transport.send_to_peers(s.build_submit_message(_page(5)))) == 0 and no_transport.stats["errors"] == 1


def test_http_bridge_posts_frames_into_the_receiver(tmp_path):
    from infomesh_b200.api.local_api import create_admin_app
    from infomesh_b200.p2p.index_submit import IndexSubmitReceiver, IndexSubmitSender
    from infomesh_b200.p2p.protocol import MessageType, encode_message

    cfg = _cfg(tmp_path, peer_acl=["only-this-peer"])
    st = _store(cfg)
    recv = IndexSubmitReceiver(cfg, st)
    c = TestClient(create_admin_app(cfg, index_submit_receiver=recv))
    frame = IndexSubmitSender(cfg).build_submit_message(_page(6))
    r = c.post("/index/submit", content=frame)
    assert r.status_code == 422 and r.json()["error"] == "peer_not_allowed"
    assert c.post("/index/submit", content=b"\x00garbage").status_code == 400
    assert c.post("/index/submit", content=encode_message(MessageType.PING, {})).json()["detail"] == "expected INDEX_SUBMIT"
    open_recv = IndexSubmitReceiver(_cfg(tmp_path), st)
    c2 = TestClient(create_admin_app(cfg, index_submit_receiver=open_recv))
    ok = c2.post("/index/submit", content=frame)
    assert ok.status_code == 200 and ok.json()["success"] and ok.json()["doc_id"] > 0
    st.close()
