"""P2P plane: envelopes, sybil, peer store, PEX, profiles, guards, bootstrap, and a live 3-node mesh on loopback
(model: reference tests/test_p2p_*.py, tests/test_sybil.py, tests/test_message_auth.py, tests/test_pex.py)."""
import asyncio
import time

import pytest

from infomesh_b200.config import Config
from infomesh_b200.p2p import message_auth as MA
from infomesh_b200.p2p import sybil
from infomesh_b200.p2p.bootstrap import (BootstrapNode, BootstrapRateLimiter, discover_bootstrap_nodes,
                                         discover_from_static, select_seed_peers)
from infomesh_b200.p2p.keys import KeyPair
from infomesh_b200.p2p.load_guard import NodeLoadGuard
from infomesh_b200.p2p.mdns import MDNSDiscovery
from infomesh_b200.p2p.network_ext import (GeoLocation, PartitionDetector, estimate_geo_distance, parse_stun_response,
                                           select_relay, sort_peers_by_proximity)
from infomesh_b200.p2p.peer_profile import BandwidthClass, PeerProfileTracker
from infomesh_b200.p2p.peer_store import PeerStore
from infomesh_b200.p2p.pex import PeerExchange
from infomesh_b200.p2p.protocol import MessageType
from infomesh_b200.p2p.replication import replica_peers
from infomesh_b200.p2p.throttle import BandwidthBucket, BandwidthThrottle
from infomesh_b200.p2p.transport import Transport, format_multiaddr, parse_multiaddr


def test_envelope_sign_verify_replay_and_tamper():
    kp = KeyPair.generate()
    reg, seen, ctr = MA.PeerKeyRegistry(), MA.NonceTracker(), MA.NonceCounter()
    reg.register(kp.peer_id, kp.public_key_bytes())
    env = MA.sign_envelope(b"hello", kp, ctr)
    assert MA.verify_envelope(env, reg, seen) == b"hello"
    with pytest.raises(MA.VerificationError, match="replayed"):
        MA.verify_envelope(env, reg, seen)
    forged = MA.SignedEnvelope(b"evil", env.peer_id, env.signature, ctr.next(), env.timestamp)
    with pytest.raises(MA.VerificationError, match="invalid signature"):
        MA.verify_envelope(forged, reg, seen)
    # the forged envelope did not burn the nonce
    env2 = MA.sign_envelope(b"again", kp, MA.NonceCounter(forged.nonce - 1))
    assert MA.verify_envelope(env2, reg, seen) == b"again"
    old = MA.sign_envelope(b"x", kp, ctr, now=time.time() - 1000)
    with pytest.raises(MA.VerificationError, match="too old"):
        MA.verify_envelope(old, reg, seen)
    with pytest.raises(MA.VerificationError, match="isolated"):
        MA.verify_envelope(MA.sign_envelope(b"y", kp, ctr), reg, seen, is_isolated_fn=lambda p: True)
    assert MA.envelope_from_dict(MA.envelope_to_dict(env)) == env


def test_pow_and_subnet_limiter():
    pub = KeyPair.generate().public_key_bytes()
    pw = sybil.generate_pow(pub, 10)
    assert sybil.verify_pow(pub, pw.nonce, 10) and sybil.leading_zero_bits(bytes.fromhex(pw.hash_hex)) >= 10
    assert not sybil.verify_pow(pub, pw.nonce, 60)
    nid = sybil.derive_node_id(pub, pw.nonce)
    assert len(nid) == 40
    v = sybil.SybilValidator(difficulty_bits=10, max_per_subnet=2)
    assert v.validate_peer(pub, pw.nonce, "10.0.0.1", nid, 3) == (True, "ok")
    assert v.validate_peer(pub, pw.nonce, "10.0.0.2", "f" * 40, 3)[1] == "node_id_mismatch"
    assert v.validate_peer(pub, pw.nonce + 1, "10.0.0.2", nid, 3)[1] in ("invalid_pow", "node_id_mismatch")
    lim = sybil.SubnetLimiter(max_per_subnet=2)
    assert lim.add("192.168.1.5", "a", 0) and lim.add("192.168.1.77", "b", 0)
    assert not lim.can_add("192.168.1.9", 0) and lim.can_add("192.168.2.9", 0) and lim.can_add("192.168.1.9", 1)
    lim.remove("192.168.1.5", "a", 0)
    assert lim.can_add("192.168.1.9", 0) and lim.total_nodes() == 1
    assert sybil.subnet_of("2001:db8:1:2::1") == "2001:db8:1::/48"


def test_peer_store_roundtrip(tmp_path):
    st = PeerStore(tmp_path)
    st.upsert("p1", "/ip4/1.2.3.4/tcp/4001/p2p/p1")
    st.upsert("p1", "/ip4/1.2.3.5/tcp/4001/p2p/p1")
    st.save_connected([("p2", "/ip4/5.6.7.8/tcp/4001/p2p/p2")])
    rec = {p.peer_id: p for p in st.load_recent()}
    assert rec["p1"].success_count == 2 and rec["p1"].multiaddr.startswith("/ip4/1.2.3.5")
    for _ in range(9):
        st.record_failure("p2")
    assert [p.peer_id for p in st.load_recent()] == ["p1"]      # 10 % success rate is hidden
    assert st.prune(max_peers=1) == 1 and st.count() == 1
    st.close()


def test_pex_rules():
    pex = PeerExchange("me")
    assert pex.check_rate_limit("a") and not pex.check_rate_limit("a")
    resp = pex.build_response([("me", "/ip4/1.1.1.1/tcp/1/p2p/me"), ("x", "/ip4/1.1.1.2/tcp/1/p2p/x"), ("bad", "tcp://x")])
    assert resp == [{"peer_id": "x", "multiaddr": "/ip4/1.1.1.2/tcp/1/p2p/x"}]
    new = pex.process_response("sender", [{"peer_id": "me", "multiaddr": "/ip4/1.1.1.1/tcp/1/p2p/me"},
                                          {"peer_id": "k", "multiaddr": "/ip4/1.1.1.3/tcp/1/p2p/k"},
                                          {"peer_id": "n", "multiaddr": "/ip4/1.1.1.4/tcp/1/p2p/n"},
                                          {"peer_id": "n", "multiaddr": "/ip4/1.1.1.4/tcp/1/p2p/n"}, "junk"], {"k"})
    assert [p.peer_id for p in new] == ["n"]


def test_peer_profiles_and_guard():
    tr = PeerProfileTracker()
    for ms in (50, 60, 70):
        tr.record("fast", ms)
    for ms in (900, 800, 1000):
        tr.record("slow", ms)
    tr.record("slow", 100, success=False)
    assert tr.get("fast").bandwidth_class == BandwidthClass.FAST and tr.get("slow").bandwidth_class == BandwidthClass.SLOW
    assert tr.get("slow").success_rate == pytest.approx(0.75)
    assert tr.rank_by_latency(["slow", "unknown", "fast"], diversity=False) == ["fast", "slow", "unknown"]
    assert 500.0 <= tr.adaptive_timeout("fast") < 700.0 and tr.adaptive_timeout("slow") == 5000.0
    assert tr.adaptive_timeout("nobody") == 2000.0
    g = NodeLoadGuard(max_queries_per_minute=3, max_concurrent=2)
    assert g.try_acquire("a") and g.try_acquire("a") and not g.try_acquire("b")
    g.release("a")
    assert g.try_acquire("b") and not g.try_acquire("c")        # qpm exhausted
    assert g.get_reject_info()["status"] == "OVERLOADED" and g.stats.rejected == 2 and g.peer_query_count("a") == 2


def test_throttle_waits():
    async def go():
        b = BandwidthBucket(0.8)                     # 100 kB/s, starts full
        t0 = time.monotonic()
        await b.acquire(100_000)
        fast = time.monotonic() - t0
        waited = await b.acquire(20_000)
        return fast, waited

    fast, waited = asyncio.run(go())
    assert fast < 0.1 and 0.1 < waited < 0.5
    th = BandwidthThrottle(0, 0)
    assert asyncio.run(th.acquire_upload(10**9)) == 0.0 and th.stats.upload_bytes == 10**9


def test_bootstrap_helpers(tmp_path):
    nodes = discover_from_static([{"addr": "/ip4/9.9.9.9/tcp/4001/p2p/x", "region": "eu"}, {"nope": 1}])
    assert len(nodes) == 1 and nodes[0].host_port == ("9.9.9.9", 4001)
    assert BootstrapNode("/dns4/boot.example/tcp/4001", "dns_srv").host_port == ("boot.example", 4001)
    res = asyncio.run(discover_bootstrap_nodes(static_nodes=[{"addr": "/ip4/9.9.9.9/tcp/4001"}], cache_dir=tmp_path,
                                               use_dns=False, use_github=False))
    assert res.addrs == ["/ip4/9.9.9.9/tcp/4001"] and res.sources_succeeded == ["static"]
    again = asyncio.run(discover_bootstrap_nodes(cache_dir=tmp_path, use_dns=False, use_github=False))
    assert again.sources_succeeded == ["cache"] and again.addrs == res.addrs
    now = time.time()
    seeds = select_seed_peers([{"peer_id": "old", "last_seen": now - 80000, "uptime": 10},
                               {"peer_id": "good", "last_seen": now, "uptime": 90000}], 1)
    assert seeds[0]["peer_id"] == "good"
    rl = BootstrapRateLimiter(max_per_minute=2)
    assert rl.allow("c") and rl.allow("c") and not rl.allow("c") and rl.tracked_clients == 1


def test_network_ext():
    seoul, tokyo, nyc = GeoLocation(latitude=37.57, longitude=126.98), GeoLocation(latitude=35.68, longitude=139.69), \
        GeoLocation(latitude=40.71, longitude=-74.0)
    assert 1100 < estimate_geo_distance(seoul, tokyo) < 1200
    assert [p for p, _ in sort_peers_by_proximity([("nyc", nyc), ("tokyo", tokyo)], seoul)] == ["tokyo", "nyc"]
    det = PartitionDetector()
    assert det.check(2, 10).is_partitioned and "Refresh routing table" in det.get_recovery_actions()
    assert not det.check(8, 10).is_partitioned and det.get_recovery_actions() == []
    assert select_relay([("a", 80.0), ("b", 20.0)]) == "b" and select_relay([]) is None
    import socket
    import struct
    txn = b"123456789012"
    xport, xip = 54321 ^ 0x2112, struct.unpack("!I", socket.inet_aton("203.0.113.7"))[0] ^ 0x2112A442
    attr = struct.pack("!HHBBHI", 0x0020, 8, 0, 1, xport, xip)
    pkt = struct.pack("!HHI", 0x0101, len(attr), 0x2112A442) + txn + attr
    assert parse_stun_response(pkt, txn) == ("203.0.113.7", 54321)
    assert parse_stun_response(pkt, b"x" * 12) is None


def test_mdns_packets():
    a, b = MDNSDiscovery("peer-a", 4001), MDNSDiscovery("peer-b", 4002)
    pkt = a.build_announce()
    assert b.observe(pkt, ("192.168.0.9", 5353)).port == 4001 and b.peer_count == 1
    assert a.observe(pkt, ("192.168.0.9", 5353)) is None           # own announcement
    assert b.observe(b"garbage", ("1.1.1.1", 1)) is None


def test_multiaddr_and_replica_choice():
    assert parse_multiaddr(format_multiaddr("10.1.2.3", 4001, "abc")) == ("10.1.2.3", 4001, "abc")
    assert parse_multiaddr("localhost:9") == ("localhost", 9, "")
    with pytest.raises(ValueError):
        parse_multiaddr("/ip4/1.2.3.4/udp/9")
    peers = [f"peer{i}" for i in range(10)]
    a = replica_peers("https://example.com/a", peers)
    assert len(a) == 3 and a == replica_peers("https://example.com/a", list(reversed(peers)))


def test_transport_signed_roundtrip_and_rejection():
    async def go():
        srv, cli = Transport(KeyPair.generate(), require_signed=True), Transport(KeyPair.generate())
        anon = Transport(None)

        async def echo(payload, peer):
            return MessageType.PONG, {"echo": payload["x"], "from": peer.peer_id}

        srv.register(MessageType.PING, echo)
        port = await srv.listen("127.0.0.1", 0)
        ok = await cli.request(("127.0.0.1", port), MessageType.PING, {"x": 7})
        bad = await anon.request(("127.0.0.1", port), MessageType.PING, {"x": 7})
        unk = await cli.request(("127.0.0.1", port), MessageType.SEARCH_REQUEST, {})
        await srv.close()
        return ok, bad, unk, cli.peer_id

    ok, bad, unk, cid = asyncio.run(go())
    assert ok == (MessageType.PONG, {"echo": 7, "from": cid})
    assert bad[0] == MessageType.ERROR and "unsigned" in bad[1]["error"]
    assert unk[0] == MessageType.ERROR


def _mk_node(tmp_path, name, boot=None, **kw):
    from infomesh_b200.p2p.node import InfoMeshNode

    from dataclasses import replace

    base = Config()
    # never "default": no network discovery in tests
    cfg = replace(base, node=replace(base.node, data_dir=tmp_path / name, listen_address="127.0.0.1", listen_port=0),
                  network=replace(base.network, bootstrap_nodes=list(boot or ["/ip4/127.0.0.1/tcp/1"]),
                                  bootstrap_dns=False, bootstrap_github=False))
    return InfoMeshNode(cfg, pow_difficulty=6, enable_mdns=False, **kw)


def test_three_node_mesh_search_publish_replicate(tmp_path):
    stored = []

    async def search_b(query, limit):
        return [{"url": "https://b.example/doc", "title": "B doc", "snippet": f"about {query}", "score": 0.9,
                 "peer_id": "", "doc_id": 5}]

    async def store_c(**kw):
        stored.append(kw)
        return True

    a = _mk_node(tmp_path, "a")
    a.start()
    boot = a.listen_addrs
    b = _mk_node(tmp_path, "b", boot, local_search_fn=search_b)
    c = _mk_node(tmp_path, "c", boot, store_fn=store_c)
    b.start()
    c.start()
    try:
        assert b.peer_id in a.get_connected_peers() and c.peer_id in a.get_connected_peers()
        # c learns about b through the iterative self-lookup during bootstrap
        assert b.peer_id in c.get_connected_peers()

        async def scenario():
            n = await b.publish_document_to_network(5, "https://b.example/doc", "B doc", "blackwell tensor memory kernels", 0.9)
            hits = await c.search_network("blackwell kernels", ["blackwell", "kernels"], 5)
            raw = await a.search_network("anything", ["zzzunknownkeyword"], 5)      # falls back to connected peers
            import hashlib
            text = "replica body text"
            reps = await a.replicate_document(1, "https://a.example/x", "t", text, hashlib.sha256(text.encode()).hexdigest())
            return n, hits, raw, reps

        n, hits, raw, reps = asyncio.run(scenario())
        assert n >= 2
        assert hits and hits[0]["url"] == "https://b.example/doc" and hits[0]["peer_id"] == b.peer_id
        assert raw and raw[0]["peer_id"] == b.peer_id
        assert reps == 1 and len(stored) == 1 and stored[0]["url"] == "https://a.example/x"
        added = a.run_sync(a._run_pex_round())
        assert added == 0                                   # everyone already knows everyone
        a._write_status_file()
        import json
        st = json.loads((tmp_path / "a" / "p2p_status.json").read_text())
        assert st["state"] == "running" and st["peers"] == 2 and st["peer_id"] == a.peer_id
        assert a.get_info().connected_peers == 2 and a.pow_nonce is not None
    finally:
        for n_ in (c, b, a):
            n_.stop()
    assert a.state == "stopped"
    import json
    assert json.loads((tmp_path / "a" / "p2p_status.json").read_text())["state"] == "stopped"
    # restart with no reachable bootstrap: rejoin through the peer store
    b2 = _mk_node(tmp_path, "b")
    a2 = _mk_node(tmp_path, "a")
    assert a2 is not None and b2 is not None
