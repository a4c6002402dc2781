"""link graph, distributed index, snapshot, vector store (CPU fallback), wire protocol."""
from __future__ import annotations

import asyncio

import pytest

from infomesh_b200.hashing import content_hash
from infomesh_b200.index.distributed import DistributedIndex, extract_keywords
from infomesh_b200.index.link_graph import LinkGraph
from infomesh_b200.index.local_store import LocalStore
from infomesh_b200.index.snapshot import export_snapshot, import_snapshot, read_snapshot_metadata
from infomesh_b200.p2p import protocol as P


def test_link_graph_authority():
    g = LinkGraph()
    assert g.add_links("https://a.com/1", ["https://hub.org/x", "https://a.com/2", "https://b.net/"]) == 3
    assert g.add_links("https://b.net/", ["https://hub.org/y"]) == 1
    assert g.add_links("https://c.io/", ["https://hub.org/z", "https://hub.org/z"]) == 1
    auth = g.compute_domain_authority()
    assert auth["hub.org"] == 1.0 and auth["hub.org"] > auth["b.net"] > 0
    assert g.url_authority("https://hub.org/anything") == 1.0 and g.url_authority("https://nope/") == 0.0
    assert g.get_stats() == {"link_count": 5, "domain_count": 4}
    assert LinkGraph().compute_domain_authority() == {}


def test_extract_keywords():
    kws = extract_keywords("The GPU runs kernels. GPU kernels use tensor cores; the GPU is fast. a b c")
    assert kws[0] == "gpu" and "kernels" in kws[:2] and "the" not in kws and "a" not in kws
    assert len(extract_keywords(" ".join(f"w{i}" for i in range(200)))) == 50


class FakeDHT:
    def __init__(self):
        self.kv: dict[str, list[dict]] = {}

    async def publish_keyword(self, kw, ptrs):
        self.kv.setdefault(kw, []).extend(ptrs)
        return True

    async def query_keyword(self, kw):
        return self.kv.get(kw, [])


def test_distributed_index_publish_query():
    dht = FakeDHT()
    a, b = DistributedIndex(dht, "peerA"), DistributedIndex(dht, "peerB")

    async def go():
        n = await a.publish_document(1, "https://x/1", "T1", "python asyncio event loop python")
        await b.publish_batch([{"doc_id": 7, "url": "https://y/7", "title": "T7", "text": "python gil"},
                               {"doc_id": 0, "url": "https://bad", "text": "ignored"}, {"doc_id": 3, "url": "", "text": "x"}])
        return n, await a.query(["python", "asyncio"])

    n, ptrs = asyncio.run(go())
    assert n == 4 and [p.peer_id for p in ptrs] == ["peerA", "peerB"] and ptrs[0].score == 2.0
    assert a.stats.documents_published == 1 and b.stats.documents_published == 1 and a.stats.queries_performed == 1


def test_snapshot_roundtrip_and_guards(tmp_path):
    src = LocalStore()
    for i in range(5):
        src.add_document(f"https://s/{i}", f"T{i}", f"document body number {i} " * 5, f"r{i}", content_hash(f"b{i}"),
                         language="en")
    path = tmp_path / "x.infomesh-snapshot"
    st = export_snapshot(src, path)
    assert st.exported == 5 and st.file_size_bytes == path.stat().st_size
    assert read_snapshot_metadata(path)["document_count"] == 5
    dst = LocalStore()
    dst.add_document("https://s/0", "T0", "document body number 0 " * 5, "r0", content_hash("b0"))
    st2 = import_snapshot(dst, path)
    assert (st2.exported, st2.skipped, st2.total_documents) == (4, 1, 5)
    assert dst.search("number")[0].url.startswith("https://s/")
    (tmp_path / "tiny").write_bytes(b"\x00\x00")
    with pytest.raises(ValueError):
        import_snapshot(dst, tmp_path / "tiny")
    (tmp_path / "bighdr").write_bytes((2**31).to_bytes(4, "big") + b"x" * 10)
    with pytest.raises(ValueError):
        read_snapshot_metadata(tmp_path / "bighdr")


def test_vector_store_cpu_fallback(tmp_path):
    from infomesh_b200.index.vector_store import VectorStore

    vs = VectorStore(tmp_path / "vec", device="cpu", max_seq_len=32)
    vs.add_document(doc_id=1, url="https://a/1", title="GPU kernels", text="tensor cores and tensor memory " * 4)
    vs.add_document(doc_id=2, url="https://a/2", title="Cooking", text="recipes for pasta and sauce " * 4)
    vs.add_document(doc_id=1, url="https://a/1", title="GPU kernels v2", text="tensor cores and tensor memory " * 4)
    assert vs.get_stats()["document_count"] == 2 and vs.get_stats()["dimension"] == 384
    res = vs.search("GPU kernels v2. tensor cores and tensor memory", limit=5)
    assert res[0].doc_id == "1" and res[0].title == "GPU kernels v2" and -1.0 <= res[0].score <= 1.0001
    assert res[0].text_preview.startswith("tensor cores")
    assert vs.search("x", min_score=2.0) == []
    vs.delete_document(1)
    assert [r.doc_id for r in vs.search("tensor", limit=5)] == ["2"]
    vs.close()
    again = VectorStore(tmp_path / "vec", device="cpu", max_seq_len=32)
    assert again.get_stats()["document_count"] == 1


def test_wire_protocol_codec():
    msg = P.encode_message(P.MessageType.SEARCH_REQUEST, {"query": "q", "keywords": ["a"], "blob": b"\x00\x01"})
    assert int.from_bytes(msg[:4], "big") == len(msg) - 4
    kind, payload = P.decode_message(msg)
    assert kind == P.MessageType.SEARCH_REQUEST and payload["blob"] == b"\x00\x01"
    assert P.decode_message(msg[4:])[0] == P.MessageType.SEARCH_REQUEST          # bare msgpack accepted
    assert P.MessageType.SIGNED_ENVELOPE == 100 and P.MessageType.CREDIT_SYNC_EXCHANGE == 73
    assert P.keyword_to_dht_key("Python") == "/infomesh/kw/" + content_hash("python")
    assert P.url_to_dht_key("https://x") == "/infomesh/url/" + content_hash("https://x")
    for bad in (b"\x00", b"\x00\x00\x00\x05abc", P.encode_message(P.MessageType.PING, {})[:6]):
        with pytest.raises(ValueError):
            P.decode_message(bad)
    with pytest.raises(ValueError):
        P.encode_message(P.MessageType.PING, {"x": b"0" * (P.MAX_MESSAGE_SIZE + 1)})
    env = P.encode_signed_envelope({"sender": "p", "sig": b"s"})
    assert P.decode_signed_envelope(env) == {"sender": "p", "sig": b"s"} and P.decode_signed_envelope(msg) is None
    assert P.dataclass_to_payload(P.PeerPointer("p", 1, "u", 0.5))["title"] == ""
    assert P.PROTOCOL_SEARCH == "/infomesh/search/1.0.0" and len(P.ALL_PROTOCOLS) == 10


def test_commoncrawl_wet_import_and_starter_helpers(tmp_path):
    import asyncio
    import gzip

    from infomesh_b200.index import commoncrawl as CC
    from infomesh_b200.index import starter as ST
    from infomesh_b200.index.local_store import LocalStore

    def rec(kind, url, body):
        return f"WARC/1.0\r\nWARC-Type: {kind}\r\nWARC-Target-URI: {url}\r\nWARC-Date: 2026-01-01T00:00:00Z\r\nContent-Length: {len(body)}\r\n\r\n{body}\r\n\r\n"

    long_a = "Blackwell tensor cores accumulate into tensor memory.\n" + "fifth generation tensor core details " * 5
    long_b = "Kademlia routing tables keep k buckets of contacts.\n" + "iterative lookups with alpha parallel requests " * 5
    wet = rec("warcinfo", "", "software: x") + rec("conversion", "https://a.example/1", long_a) + rec("conversion", "https://a.example/2", long_b) \
        + rec("conversion", "https://a.example/dup", long_a) + rec("conversion", "https://a.example/short", "tiny")
    recs = CC.parse_wet_content(wet)
    assert [r.url for r in recs] == ["https://a.example/1", "https://a.example/2", "https://a.example/dup"]
    p = tmp_path / "sample.wet.gz"
    p.write_bytes(gzip.compress(wet.encode()))
    store = LocalStore(tmp_path / "cc.db")
    imp = CC.CommonCrawlImporter(store)
    st = asyncio.run(imp.import_wet_file(str(p)))
    assert (st.total_records, st.imported, st.skipped_duplicate) == (3, 2, 1)
    assert store.get_document_by_url("https://a.example/1").title.startswith("Blackwell tensor cores")
    ul = tmp_path / "urls.txt"
    ul.write_text("# list\nhttps://b.example/x\nhttps://b.example/y\nhttps://a.example/1\n")
    st2 = asyncio.run(imp.import_url_list(ul))
    assert (st2.total_records, st2.imported, st2.skipped_duplicate) == (3, 2, 1)
    import io
    import pytest
    with pytest.raises(ValueError, match="exceeds"):
        CC._read_binary_limited(io.BytesIO(b"x" * 100), "mem", limit=10)
    store.close()
    rel = [{"tag_name": "v2", "assets": [{"name": "other.bin"}]},
           {"tag_name": "v1", "assets": [{"name": ST.SNAPSHOT_ASSET_NAME, "browser_download_url": "https://x/s", "size": 2 ** 21, "created_at": "c"}]}]
    info = ST.pick_asset(rel)
    assert info.release_tag == "v1" and info.size_mb == 2.0 and ST.pick_asset([]) is None
    ST._write_cache(tmp_path, info)
    assert ST._read_cache(tmp_path).download_url == "https://x/s" and ST.needs_starter(3) and not ST.needs_starter(10)


def test_local_store_survives_concurrent_readers_and_a_writer(tmp_path):
    """One connection, many threads (MCP handlers + GPU index workers + crawler): reads and writes must serialise instead of
    tripping sqlite3's 'bad parameter or other API misuse' on a shared connection."""
    import threading

    from infomesh_b200.index.local_store import LocalStore

    st = LocalStore(tmp_path / "c.db")
    for i in range(50):
        st.add_document(url=f"https://e.example/{i}", title=f"T{i}", text=f"alpha beta gamma document {i} " * 5, raw_html_hash=f"r{i}", text_hash=f"t{i}")
    errors, stop = [], threading.Event()

    def reader():
        try:
            while not stop.is_set():
                for i in range(1, 40, 3):
                    d = st.get_document(i)
                    assert d is not None and d.url.endswith(str(i - 1))
                assert st.search("alpha beta", limit=5)
                assert st.get_stats()["document_count"] >= 50
                for _ in zip(range(5), st.iter_documents(batch=7)):
                    pass
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=reader) for _ in range(4)]
    for t in threads:
        t.start()
    for i in range(50, 120):
        st.add_document(url=f"https://e.example/{i}", title=f"T{i}", text=f"alpha beta delta document {i} " * 5, raw_html_hash=f"r{i}", text_hash=f"t{i}")
    stop.set()
    for t in threads:
        t.join()
    assert not errors, errors[:2]
    assert st.get_stats()["document_count"] == 120
    st.close()
