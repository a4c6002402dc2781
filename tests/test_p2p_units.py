"""p2p/peer_profile.py, load_guard.py, throttle.py, pex.py, peer_store.py, replication.py, protocol.py codec edges."""
import asyncio
import hashlib
import time
from unittest.mock import patch

import msgpack
import pytest

from infomesh_b200.p2p import peer_profile as PP
from infomesh_b200.p2p import protocol as PR
from infomesh_b200.p2p.load_guard import NodeLoadGuard
from infomesh_b200.p2p.peer_store import PeerStore
from infomesh_b200.p2p.pex import PEX_MAX_PEERS, PeerExchange
from infomesh_b200.p2p.replication import Replicator, replica_peers
from infomesh_b200.p2p.throttle import BandwidthBucket, BandwidthThrottle


# ------------------------------------------------------------------ peer profiles
def test_profile_ema_percentile_success_rate_and_class():
    t = PP.PeerProfileTracker()
    p = t.record("a", 100.0)
    assert p.avg_latency_ms == 100.0 and p.bandwidth_class == PP.BandwidthClass.UNKNOWN
    t.record("a", 200.0)
    p = t.record("a", 50.0, success=False)                       # failures do not touch latency
    assert p.avg_latency_ms == pytest.approx(0.3 * 200 + 0.7 * 100) and p.success_rate == pytest.approx(2 / 3)
    assert p.bandwidth_class == PP.BandwidthClass.MEDIUM and p.total_interactions == 3
    assert PP._percentile([1, 2, 3, 4], 50) == 2.5 and PP._percentile([], 95) == 0.0
    assert [PP._classify_bandwidth(x) for x in (50, 100, 499, 500)] == [PP.BandwidthClass.FAST, PP.BandwidthClass.MEDIUM,
                                                                         PP.BandwidthClass.MEDIUM, PP.BandwidthClass.SLOW]


def test_rank_by_latency_and_adaptive_timeout():
    t = PP.PeerProfileTracker()
    for pid, ms in (("fast", 20), ("mid", 150), ("slow", 900)):
        for _ in range(3):
            t.record(pid, ms)
    assert t.rank_by_latency(["slow", "unknown", "fast", "mid"], diversity=False) == ["fast", "mid", "slow", "unknown"]
    with patch("random.random", return_value=0.99):
        assert t.rank_by_latency(["slow", "unknown", "fast", "mid"]) == ["fast", "mid", "slow", "unknown"]
    with patch("random.random", return_value=0.0):
        assert t.rank_by_latency(["slow", "unknown", "fast", "mid"])[:2] == ["fast", "mid"]
    assert t.adaptive_timeout("nobody") == 2000.0 and t.adaptive_timeout("fast") == 500.0
    assert t.adaptive_timeout("slow") == 5000.0 and t.adaptive_timeout("mid") == pytest.approx(2000 * 150 / 200)


def test_profile_pruning():
    t = PP.PeerProfileTracker()
    t.record("old", 10)
    t.get("old").last_seen = time.time() - 2 * PP.STALE_TIMEOUT
    t.record("new", 10)
    assert t.prune_stale() == 1 and t.known_peers == 1 and t.get("old") is None
    t.reset()
    assert t.known_peers == 0 and t.get_or_default("x").peer_id == "x"


# ------------------------------------------------------------------ load guard
def test_load_guard_concurrency_and_rate_caps():
    g = NodeLoadGuard(max_queries_per_minute=3, max_concurrent=2)
    assert g.try_acquire("a") and g.try_acquire("a") and not g.try_acquire("b")      # concurrency cap
    g.release("a")
    assert g.try_acquire("b") and g.stats.queries_this_minute == 3
    g.release("a"), g.release("b")
    assert not g.try_acquire("c") and g.is_overloaded                                  # per-minute cap
    info = g.get_reject_info()
    assert info["status"] == "OVERLOADED" and info["qpm"] == 3 and g.stats.rejected == 2 and g.stats.accepted == 3
    assert g.peer_query_count("a") == 2 and g.peer_query_count("zz") == 0
    g.reset()
    assert g.try_acquire() and not g.is_overloaded


# ------------------------------------------------------------------ throttle
def test_bandwidth_bucket_waits_only_beyond_burst():
    async def go():
        b = BandwidthBucket(8.0)                                   # 1 MB/s, 1 MB burst
        assert b.rate_bytes_per_sec == 1_000_000
        assert await b.acquire(0) == 0.0 and await b.acquire(900_000) == 0.0
        t0 = time.monotonic()
        waited = await b.acquire(300_000)                          # 200 KB short -> ~0.2 s
        return waited, time.monotonic() - t0

    waited, wall = asyncio.run(go())
    assert 0.1 < waited < 0.5 and wall >= waited * 0.8


def test_throttle_stats_and_unlimited_direction():
    async def go():
        t = BandwidthThrottle(upload_mbps=0, download_mbps=0.8)    # 100 KB/s down, unlimited up
        assert await t.acquire_upload(10_000_000) == 0.0
        w = await t.acquire_download(120_000)
        return t.stats, w

    st, w = asyncio.run(go())
    assert st.upload_bytes == 10_000_000 and st.upload_waits == 0 and st.download_waits == 1 and w > 0.1


# ------------------------------------------------------------------ PEX
def test_pex_build_and_process_responses():
    px = PeerExchange("me")
    good = "/ip4/10.0.0.1/tcp/4001/p2p/QmA"
    resp = px.build_response([("me", good), ("a", good), ("b", "tcp://bad"), ("c", "/ip6/::1/tcp/1/p2p/QmC")] + [(f"x{i}", good) for i in range(20)])
    assert [r["peer_id"] for r in resp[:2]] == ["a", "c"] and len(resp) == PEX_MAX_PEERS
    got = px.process_response("sender", [{"peer_id": "n1", "multiaddr": good}, {"peer_id": "n1", "multiaddr": good}, {"peer_id": "me", "multiaddr": good},
                                         {"peer_id": "sender", "multiaddr": good}, {"peer_id": "known", "multiaddr": good}, {"peer_id": "bad", "multiaddr": 5},
                                         "junk", {"peer_id": "", "multiaddr": good}], known_peers={"known"})
    assert [p.peer_id for p in got] == ["n1"]


def test_pex_rate_limit_window():
    px = PeerExchange("me")
    assert px.check_rate_limit("p") and not px.check_rate_limit("p") and px.check_rate_limit("q")
    px._last_request["old"] = time.time() - 10_000
    px.cleanup_rate_limits()
    assert "old" not in px._last_request and "p" in px._last_request


# ------------------------------------------------------------------ peer store
def test_peer_store_success_rate_ordering_and_prune(tmp_path):
    ps = PeerStore(tmp_path / "peers.db")
    ps.save_connected([("a", "/ip4/1.1.1.1/tcp/1/p2p/a"), ("b", "/ip4/2.2.2.2/tcp/1/p2p/b")])
    ps.upsert("a", "/ip4/1.1.1.1/tcp/2/p2p/a")
    for _ in range(3):
        ps.record_failure("b")
    rec = {p.peer_id: p for p in ps.load_recent()}
    assert rec["a"].multiaddr.endswith("/tcp/2/p2p/a") and rec["a"].success_rate == 1.0 and rec["b"].success_rate < 0.5
    assert ps.load_recent(limit=1)[0].peer_id == "a" and ps.count() == 2
    ps.remove("a")
    assert ps.count() == 1 and ps.prune(max_age_hours=0) == 1 and ps.count() == 0
    ps.close()


# ------------------------------------------------------------------ replication
def test_replica_peers_are_deterministic_and_url_dependent():
    peers = [f"peer{i}" for i in range(10)]
    a = replica_peers("https://ex.org/a", peers)
    assert a == replica_peers("https://ex.org/a", list(reversed(peers))) and len(a) == 3 and len(set(a)) == 3
    assert any(replica_peers(f"https://ex.org/{i}", peers) != a for i in range(5))
    assert replica_peers("u", ["only"], 3) == ["only"]


def test_replicator_counts_successes_failures_and_validates_inbound():
    sent = []

    async def send(pid, kind, payload, timeout):
        sent.append((pid, payload["replica_index"]))
        if pid == "down":
            raise OSError("unreachable")
        return PR.MessageType.REPLICATE_RESPONSE, {"success": pid != "refuses"}

    rep = Replicator(send, lambda: ["me", "ok", "down", "refuses"], "me")
    n = asyncio.run(rep.replicate_document(1, "https://ex.org/x", "T", "body", "h"))
    assert n == 1 and rep.stats.replicas_sent == 1 and rep.stats.replicas_failed == 2 and rep.stats.documents_replicated == 1
    assert sorted(i for _, i in sent) == [0, 1, 2] and all(p != "me" for p, _ in sent)
    assert asyncio.run(Replicator(send, lambda: ["me"], "me").replicate_document(1, "u", "t", "x", "h")) == 0

    stored = []

    async def store(**kw):
        stored.append(kw)
        return True

    body = "replicated text"
    ok = asyncio.run(rep.handle_replicate_request({"url": "u", "text": body, "text_hash": hashlib.sha256(body.encode()).hexdigest(), "title": "T"}, store))
    assert ok[1]["success"] and stored[0]["title"] == "T" and rep.stats.replicas_received == 1
    assert asyncio.run(rep.handle_replicate_request({"url": "u", "text": body, "text_hash": "bad"}, store))[1]["error"] == "hash_mismatch"
    assert asyncio.run(rep.handle_replicate_request({"url": "", "text": body}, store))[1]["error"] == "empty"

    async def broken(**kw):
        raise RuntimeError("disk full")

    assert not asyncio.run(rep.handle_replicate_request({"url": "u", "text": body}, broken))[1]["success"]


# ------------------------------------------------------------------ wire codec
def test_codec_roundtrip_and_frame_validation():
    frame = PR.encode_message(PR.MessageType.PING, {"n": 1, "b": b"\x00\x01"})
    assert int.from_bytes(frame[:4], "big") == len(frame) - 4
    assert PR.decode_message(frame) == (PR.MessageType.PING, {"n": 1, "b": b"\x00\x01"})
    assert PR.read_frame_length(frame[:4]) == len(frame) - 4
    for bad in (b"\x00\x00\x00\x00", (PR.MAX_MESSAGE_SIZE + 1).to_bytes(4, "big"), b"\x00\x01"):
        with pytest.raises(ValueError):
            PR.read_frame_length(bad)
    with pytest.raises(ValueError):
        PR.encode_message(PR.MessageType.PING, {"blob": b"x" * (PR.MAX_MESSAGE_SIZE + 10)})


def test_decode_rejects_malformed_and_unknown_types():
    body = msgpack.packb({"type": 9999, "payload": {}})
    with pytest.raises(ValueError, match="Unknown message type"):
        PR.decode_message(len(body).to_bytes(4, "big") + body)
    body = msgpack.packb({"payload": {}})
    with pytest.raises(ValueError, match="missing"):
        PR.decode_message(len(body).to_bytes(4, "big") + body)
    with pytest.raises(ValueError):
        PR.decode_message(b"\x00")
    with pytest.raises(ValueError):
        PR.decode_message(b"\x00\x00\x00\x05\xc1\xc1\xc1\xc1\xc1")
    huge = msgpack.packb({"type": 0, "payload": {"a": list(range(70_000))}})
    with pytest.raises(ValueError):
        PR.decode_message(len(huge).to_bytes(4, "big") + huge)            # array longer than the safe-unpack limit


def test_dht_keys_and_envelope_helpers():
    assert PR.keyword_to_dht_key("Python") == PR.keyword_to_dht_key("python") and PR.keyword_to_dht_key("a").startswith("/infomesh/kw/")
    assert PR.url_to_dht_key("https://a") != PR.url_to_dht_key("https://A") and len(PR.ALL_PROTOCOLS) == 10
    env = {"payload": b"p", "peer_id": "x", "signature": b"s", "nonce": 1, "timestamp": 2.0}
    assert PR.decode_signed_envelope(PR.encode_signed_envelope(env)) == env
    assert PR.decode_signed_envelope(PR.encode_message(PR.MessageType.PING, {})) is None
    assert PR.dataclass_to_payload(PR.SearchRequest("q", ["k"], 5, "r", 1.0)) == {"query": "q", "keywords": ["k"], "limit": 5, "request_id": "r", "timestamp": 1.0}
