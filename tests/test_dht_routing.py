"""p2p/dht.py over a dict-backed fake KadDHT, p2p/routing.py with in-memory peers, p2p/kademlia.py routing table
(reference shape: hand-written fakes injected through constructors, tests/test_dht.py / test_routing.py)."""
import asyncio
import time

import msgpack

from infomesh_b200.p2p import dht as DH
from infomesh_b200.p2p import kademlia as K
from infomesh_b200.p2p.load_guard import NodeLoadGuard
from infomesh_b200.p2p.protocol import MessageType, keyword_to_dht_key
from infomesh_b200.p2p.routing import QueryRouter, _payload_float, _payload_int, _payload_str
from infomesh_b200.p2p.sybil import SubnetLimiter


class FakeKad:
    def __init__(self, fail=False):
        self.kv, self.fail = {}, fail

    async def put_value(self, key, value):
        if self.fail:
            raise OSError("network down")
        self.kv[key] = value

    async def get_value(self, key):
        if self.fail:
            raise OSError("network down")
        return self.kv.get(key)


def run(coro):
    return asyncio.run(coro)


def ptr(peer, doc, score=1.0):
    return {"peer_id": peer, "doc_id": doc, "url": f"https://x/{doc}", "score": score, "title": "t"}


def test_publish_merges_new_first_and_dedups():
    d = DH.InfoMeshDHT(FakeKad(), "me")
    assert run(d.publish_keyword("Python", [ptr("a", 1), ptr("b", 2)]))
    assert run(d.publish_keyword("python", [ptr("a", 1, 9.0), ptr("c", 3)]))          # same key: case-folded
    got = run(d.query_keyword("PYTHON"))
    assert [(p["peer_id"], p["doc_id"]) for p in got] == [("a", 1), ("c", 3), ("b", 2)] and got[0]["score"] == 9.0
    assert d.stats.keys_published == 2 and d.stats.puts_performed == 2


def test_publish_caps_pointers_and_drops_malformed():
    d = DH.InfoMeshDHT(FakeKad(), "me")
    many = [ptr("p", i) for i in range(150)] + [{"peer_id": 5, "doc_id": 1}, "junk", {"peer_id": "x", "doc_id": True}]
    run(d.publish_keyword("k", many))
    got = run(d.query_keyword("k"))
    assert len(got) == DH.MAX_POINTERS_PER_KEYWORD and all(isinstance(p["doc_id"], int) for p in got)


def test_publish_rate_limit_per_keyword():
    d = DH.InfoMeshDHT(FakeKad(), "me")
    for i in range(DH.MAX_PUBLISHES_PER_KEYWORD_HR):
        assert run(d.publish_keyword("hot", [ptr("a", i)]))
    assert not run(d.publish_keyword("hot", [ptr("a", 99)]))
    assert run(d.publish_keyword("other", [ptr("a", 1)]))


def test_backend_failures_degrade_to_false_and_empty():
    d = DH.InfoMeshDHT(FakeKad(fail=True), "me")
    assert not run(d.put("k", b"v")) and run(d.get("k")) is None
    assert run(d.query_keyword("k")) == [] and not run(d.publish_keyword("k", [ptr("a", 1)]))


def test_query_ignores_garbage_values():
    kad = FakeKad()
    d = DH.InfoMeshDHT(kad, "me")
    kad.kv[keyword_to_dht_key("bad")] = b"\xff\xff not msgpack"
    kad.kv[keyword_to_dht_key("list")] = msgpack.packb([1, 2, 3])
    kad.kv[keyword_to_dht_key("wrong")] = msgpack.packb({"pointers": "nope"})
    assert run(d.query_keyword("bad")) == [] and run(d.query_keyword("list")) == [] and run(d.query_keyword("wrong")) == []


def test_crawl_lock_contention_ttl_and_release():
    kad = FakeKad()
    a, b = DH.InfoMeshDHT(kad, "A"), DH.InfoMeshDHT(kad, "B")
    url = "https://ex.org/page"
    assert run(a.acquire_crawl_lock(url)) and not run(a.acquire_crawl_lock(url))     # exclusive, also against the holder's second attempt
    assert not run(b.acquire_crawl_lock(url))
    assert run(b.acquire_crawl_lock(url, ttl_seconds=0))                              # expired from B's point of view
    assert run(b.release_crawl_lock(url)) and run(a.acquire_crawl_lock(url))
    assert a.stats.locks_acquired == 2 and b.stats.locks_released == 1


def test_attestation_roundtrip():
    d = DH.InfoMeshDHT(FakeKad(), "me")
    assert run(d.publish_attestation("https://u", "raw", "text", b"sig"))
    att = run(d.get_attestation("https://u"))
    assert att["raw_hash"] == "raw" and att["peer_id"] == "me" and att["signature"] == b"sig"
    assert run(d.get_attestation("https://none")) is None


# ------------------------------------------------------------------ routing
class FakeDht:
    def __init__(self, table):
        self.table = table

    async def query_keyword(self, kw):
        v = self.table.get(kw, [])
        if isinstance(v, Exception):
            raise v
        return v


def make_router(table, replies, **kw):
    sent = []

    async def send(pid, kind, payload, timeout_s):
        sent.append((pid, kind, payload))
        r = replies.get(pid)
        if isinstance(r, Exception):
            raise r
        if r == "slow":
            await asyncio.sleep(5)
        return r

    return QueryRouter(send, FakeDht(table), "me", **kw), sent


def resp(*rows):
    return MessageType.SEARCH_RESPONSE, {"results": list(rows)}


def test_route_query_merges_best_score_per_url_and_skips_self():
    table = {"python": [ptr("p1", 1, 2.0), ptr("me", 9), ptr("p2", 2, 1.0)], "rust": RuntimeError("dht hiccup")}
    replies = {"p1": resp({"url": "u1", "title": "A", "snippet": "s", "score": 0.4, "doc_id": 1}, {"url": "", "score": 9}, "junk"),
               "p2": resp({"url": "u1", "title": "A2", "snippet": "s", "score": 0.9, "doc_id": 7}, {"url": "u2", "score": float("nan"), "doc_id": "2"})}     # numbers may arrive as text; a float id would not be truncated
    r, sent = make_router(table, replies)
    out = run(r.route_query("python rust", ["python", "rust"], limit=5))
    assert [(x.url, x.peer_id, x.score) for x in out] == [("u1", "p2", 0.9), ("u2", "p2", 0.0)] and out[1].doc_id == 2
    assert {s[0] for s in sent} == {"p1", "p2"} and sent[0][2]["query"] == "python rust"
    assert r.stats.peers_contacted == 2 and r.stats.peers_responded == 2 and r.stats.avg_response_ms >= 0


def test_route_query_falls_back_to_connected_peers_then_local_only():
    r, sent = make_router({}, {"c1": resp({"url": "u", "score": 1.0})}, connected_peers=lambda: ["c1", "me", ""])
    assert [x.url for x in run(r.route_query("q", ["none"]))] == ["u"] and [s[0] for s in sent] == ["c1"]
    r2, sent2 = make_router({}, {})
    assert run(r2.route_query("q", ["none"])) == [] and r2.stats.queries_local_only == 1 and sent2 == []
    assert run(r2.route_query("q", ["none"], limit=0)) == []


def test_route_query_counts_failures_and_respects_fanout():
    table = {"k": [ptr(f"p{i}", i, 10 - i) for i in range(8)]}
    replies = {"p0": OSError("refused"), "p1": "slow", "p2": (MessageType.ERROR, {"reason": "busy"}), "p3": None}
    r, sent = make_router(table, replies, max_fanout=4, timeout_ms=150)
    t0 = time.monotonic()
    assert run(r.route_query("q", ["k"])) == []
    assert time.monotonic() - t0 < 3 and len(sent) == 4
    assert r.stats.peers_timed_out == 2 and r.stats.peers_responded == 0


def test_handle_search_request_limits_and_load_shedding():
    r, _ = make_router({}, {}, load_guard=NodeLoadGuard(max_queries_per_minute=2, max_concurrent=5))
    calls = []

    async def local(q, n):
        calls.append((q, n))
        return [{"url": f"u{i}"} for i in range(n + 5)]

    kind, payload = run(r.handle_search_request({"query": "hello", "limit": 3, "request_id": "r1"}, local, requester="x"))
    assert kind == MessageType.SEARCH_RESPONSE and len(payload["results"]) == 3 and payload["request_id"] == "r1" and payload["peer_id"] == "me"
    kind, payload = run(r.handle_search_request({"query": "   ", "limit": 10_000}, local, requester="x"))
    assert payload["results"] == [] and calls == [("hello", 3)]
    kind, payload = run(r.handle_search_request({"query": "again"}, local, requester="x"))
    assert kind == MessageType.ERROR and payload


def test_payload_coercion_helpers():
    assert _payload_str(5, default="d") == "d" and _payload_str("x") == "x"
    # the reference's contract (p2p/routing.py:415-436): integer strings count, floats are never truncated into ids, NaN / inf are dropped
    assert _payload_int(True, default=7) == 7 and _payload_int(3.9) == 0 and _payload_int(float("inf"), default=1) == 1 and _payload_int("4") == 4
    assert _payload_int("x", default=9) == 9 and _payload_int(" 12 ") == 12 and _payload_int([1]) == 0
    assert _payload_float("1.5", default=2.0) == 1.5 and _payload_float(float("nan")) == 0.0 and _payload_float(2) == 2.0 and _payload_float("inf", default=3.0) == 3.0


# ------------------------------------------------------------------ kademlia routing table
def pid(i):
    return f"{i:040x}"


def test_key_and_node_ids():
    assert K.key_id("a") == K.key_id(b"a") and K.key_id("a") != K.key_id("b") and K.key_id("a") < 1 << 160
    assert K.node_int(pid(255)) == 255 and K.node_int("not-hex") == K.key_id("not-hex")


def test_routing_table_buckets_lru_and_capacity():
    rt = K.RoutingTable(pid(0), k=2)
    assert rt.bucket_index(pid(1)) == 0 and rt.bucket_index(pid(2)) == 1 and rt.bucket_index(pid(0xFF)) == 7
    assert not rt.add(K.Contact(pid(0), "h", 1)) and not rt.add(K.Contact("", "h", 1))
    assert rt.add(K.Contact(pid(4), "h", 1)) and rt.add(K.Contact(pid(5), "h", 2)) and not rt.add(K.Contact(pid(6), "h", 3))   # bucket 2 full
    assert rt.add(K.Contact(pid(4), "h2", 9)) and list(rt.buckets[2]) == [pid(5), pid(4)] and rt.get(pid(4)).host == "h2"
    rt.remove(pid(5))
    assert rt.add(K.Contact(pid(6), "h", 3)) and len(rt) == 2 and rt.get(pid(5)) is None


def test_routing_table_closest_by_xor_and_subnet_quota():
    rt = K.RoutingTable(pid(0), subnet_limiter=SubnetLimiter(max_per_subnet=1))
    for i, host in ((1, "10.0.0.1"), (2, "10.0.1.1"), (8, "10.0.2.1"), (9, "10.0.2.2"), (12, "node.example.org")):
        rt.add(K.Contact(pid(i), host, 4000 + i))
    assert rt.get(pid(9)) is None                                        # same /24 in the same bucket as 8
    assert [K.node_int(c.peer_id) for c in rt.closest(9, 3)] == [8, 12, 1]
    rt.remove(pid(8))
    assert rt.add(K.Contact(pid(9), "10.0.2.2", 4009)) and rt.get(pid(12)).to_wire()["host"] == "node.example.org"
