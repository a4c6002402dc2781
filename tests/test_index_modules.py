"""index/local_store.py (CRUD, FTS sync, stats, listeners), link_graph.py, distributed.py, snapshot.py, vector_store.py."""
import asyncio

import pytest

from infomesh_b200.hashing import content_hash
from infomesh_b200.index.distributed import DistributedIndex, extract_keywords
from infomesh_b200.index.link_graph import LinkGraph
from infomesh_b200.index.local_store import LocalStore
from infomesh_b200.index.snapshot import export_snapshot, import_snapshot, read_snapshot_metadata


def add(store, url, title, text, **kw):
    return store.add_document(url, title, text, content_hash("raw" + url), content_hash(text), **kw)


# ------------------------------------------------------------------ LocalStore
def test_store_rejects_unknown_tokenizer_and_duplicates():
    with pytest.raises(ValueError):
        LocalStore(None, tokenizer="porter; DROP TABLE documents")
    s = LocalStore(None)
    a = add(s, "https://a.com/1", "One", "alpha beta gamma delta")
    assert isinstance(a, int) and add(s, "https://a.com/1", "again", "different text entirely") is None       # same URL
    assert add(s, "https://a.com/2", "Copy", "alpha beta gamma delta") is None                               # same text hash
    assert s.get_stats()["document_count"] == 1


def test_store_update_keeps_fts_in_sync_and_delete_removes_hits():
    s = LocalStore(None)
    add(s, "https://a.com/1", "Original title", "the quick brown fox")
    assert [r.url for r in s.search("fox")] == ["https://a.com/1"]
    assert s.update_document("https://a.com/1", text="a lazy dog sleeps", text_hash=content_hash("a lazy dog sleeps"), title="New title")
    assert s.search("fox") == [] and s.search("dog")[0].title == "New title"
    assert not s.update_document("https://a.com/1") and not s.update_document("https://missing", title="x")
    doc = s.get_document_by_url("https://a.com/1")
    assert s.update_document(doc.url, stale_count=2, change_frequency=0.4, etag='"e"') and s.get_document(doc.doc_id).stale_count == 2
    assert s.soft_delete("https://a.com/1") and not s.soft_delete("https://a.com/1") and s.search("dog") == []
    assert not s.delete_document(12345)


def test_store_search_limits_bad_syntax_and_suggest():
    s = LocalStore(None)
    for i in range(30):
        add(s, f"https://a.com/{i}", f"Python tip {i}", f"python tip number {i} about decorators and generators {i}")
    assert len(s.search("python", limit=5)) == 5 and len(s.search("python", limit=5, offset=28)) == 2
    assert len(s.search("python", limit=100000)) == 30 and s.search('"unbalanced') == []                   # FTS error -> empty
    assert all(r.score > 0 for r in s.search("decorators"))
    sug = s.suggest("Pyth", limit=3)
    assert len(sug) <= 3 and all("python" in t.lower() for t in sug) and s.suggest("%%%") == s.suggest("")


def test_store_listeners_stats_domains_and_iteration():
    s = LocalStore(None)
    events = []
    s.add_listener(lambda ev, payload: events.append((ev, payload.get("url") or payload.get("doc_id"))))
    s.add_listener(lambda ev, payload: 1 / 0)                                                               # a broken listener is ignored
    d1 = add(s, "https://a.com/1", "A1", "text one about apples", language="en", js_required=True)
    add(s, "https://a.com/2", "A2", "text two about bananas")
    add(s, "https://b.org/1", "B1", "text three about cherries")
    s.update_document("https://a.com/2", title="A2b")
    s.delete_document(d1)
    assert [e[0] for e in events] == ["add", "add", "add", "update", "delete"]
    assert sorted(s.get_top_domains()) == [("a.com", 1), ("b.org", 1)] and s.get_domain_count() == 2
    assert [d.url for d in s.iter_documents(batch=1)] == ["https://a.com/2", "https://b.org/1"]
    assert {d["url"] for d in s.export_documents()} == {"https://a.com/2", "https://b.org/1"}
    assert len(s.get_documents_for_publish(limit=1)) == 1 and isinstance(s.get_recrawl_candidates(), list)
    s.optimize()


def test_store_compression_roundtrip(tmp_path):
    with LocalStore(tmp_path / "i.db", compression_enabled=True) as s:
        text = "compressible text " * 400
        did = add(s, "https://a.com/z", "Z", text)
        assert s.get_document(did).text == text
        st = s.get_compression_stats()
        assert st and (st.get("compressed_bytes", 1) or 1) < len(text)
    with LocalStore(tmp_path / "i.db") as again:
        assert again.get_document_by_url("https://a.com/z").title == "Z"


# ------------------------------------------------------------------ link graph
def test_link_graph_authority_flows_to_linked_domains():
    g = LinkGraph()
    assert g.compute_domain_authority() == {} and g.url_authority("https://x.org/") == 0.0
    n = g.add_links("https://blog.a.com/post", ["https://docs.hub.org/x", "https://docs.hub.org/y", "https://blog.a.com/other",
                                                 "https://blog.a.com/post", "not a url"])
    assert n == 3
    g.add_links("https://b.net/p", ["https://docs.hub.org/x"])
    g.add_links("https://c.io/p", ["https://docs.hub.org/z", "https://b.net/p"])
    auth = g.compute_domain_authority(use_gpu=False)
    assert max(auth, key=auth.get) == "docs.hub.org" and auth["docs.hub.org"] == 1.0 and 0 < auth["c.io"] < auth["b.net"] < 1
    assert g.url_authority("https://docs.hub.org/anything") == 1.0 and g.domain_authority("unknown.tld") == 0.0
    st = g.get_stats()
    assert st["link_count"] == 6 and st["domain_count"] >= 4
    g.close()


# ------------------------------------------------------------------ distributed index
class FakeDht:
    def __init__(self):
        self.kw = {}

    async def publish_keyword(self, kw, ptrs):
        self.kw.setdefault(kw, []).extend(ptrs)
        return kw != "refused"

    async def query_keyword(self, kw):
        return self.kw.get(kw, [])


def test_extract_keywords_frequency_stopwords_and_order():
    kws = extract_keywords("The python python asyncio loop and the loop of python is a loop", max_keywords=3)
    assert kws == ["python", "loop", "asyncio"] and extract_keywords("a an of the") == []


def test_distributed_index_publish_and_query_aggregates_scores():
    dht = FakeDht()
    di = DistributedIndex(dht, "me")
    n = asyncio.run(di.publish_batch([
        {"doc_id": 1, "url": "https://a/1", "title": "A", "text": "python asyncio event loop", "score": 1.0},
        {"doc_id": 2, "url": "https://a/2", "title": "B", "text": "python typing generics", "score": 0.5},
        {"doc_id": 0, "url": "https://bad", "text": "ignored: bad id"}, {"doc_id": 3, "url": "", "text": "ignored: no url"},
        {"doc_id": 4, "url": "https://a/4", "text": ""}]))
    assert n == len(dht.kw) and di.stats.documents_published == 2 and len(dht.kw["python"]) == 2
    ptrs = asyncio.run(di.query(["python", "asyncio", "missing"]))
    assert [(p.doc_id, p.score) for p in ptrs] == [(1, 2.0), (2, 0.5)] and ptrs[0].peer_id == "me" and di.stats.pointers_found == 2
    assert asyncio.run(di.publish_document(9, "https://a/9", "T", "refused refused refused")) == 0


# ------------------------------------------------------------------ snapshots
def test_snapshot_roundtrip_metadata_and_corruption(tmp_path):
    src = LocalStore(None)
    for i in range(12):
        add(src, f"https://s.org/{i}", f"Doc {i}", f"snapshot body number {i} with some words", language="en")
    path = tmp_path / "snap.imsnap"
    st = export_snapshot(src, path)
    assert st.total_documents == 12 and path.stat().st_size > 0
    meta = read_snapshot_metadata(path)
    assert meta["document_count"] == 12 if "document_count" in meta else meta
    dst = LocalStore(None)
    add(dst, "https://s.org/3", "pre-existing", "already here so it must be skipped")
    res = import_snapshot(dst, path)
    assert dst.get_stats()["document_count"] == 12 and res.total_documents == 12
    assert dst.search("snapshot")[0].url.startswith("https://s.org/")
    bad = tmp_path / "bad.imsnap"
    bad.write_bytes(b"not a snapshot")
    with pytest.raises(Exception):
        import_snapshot(LocalStore(None), bad)


# ------------------------------------------------------------------ vector store (CPU path)
def test_vector_store_semantic_search_delete_and_persistence(tmp_path):
    from infomesh_b200.index.vector_store import VectorStore

    vs = VectorStore(tmp_path / "vec", device="cpu") if "device" in VectorStore.__init__.__code__.co_varnames else VectorStore(tmp_path / "vec")
    vs.add_documents([{"doc_id": 1, "url": "https://a/1", "title": "Cats", "text": "cats purr and chase mice around the house"},
                      {"doc_id": 2, "url": "https://a/2", "title": "GPUs", "text": "tensor cores multiply matrices on the gpu"}])
    vs.add_document(doc_id=3, url="https://a/3", title="Dogs", text="dogs bark and fetch sticks in the park")
    hits = vs.search("cats purr and chase mice around the house", limit=2)
    assert str(hits[0].doc_id) == "1" and hits[0].score > hits[-1].score - 1e-6 and len(vs.embed(["x"])[0]) > 8
    vs.delete_document(1)
    assert all(str(h.doc_id) != "1" for h in vs.search("cats purr and chase mice", limit=3))
    assert vs.get_stats()["document_count"] == 2
    vs.close()
    again = VectorStore(tmp_path / "vec", device="cpu") if "device" in VectorStore.__init__.__code__.co_varnames else VectorStore(tmp_path / "vec")
    assert again.get_stats()["document_count"] == 2 and again.search("dogs bark and fetch sticks in the park", limit=1)[0].doc_id in (3, "3")
    again.close()
