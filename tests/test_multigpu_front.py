"""MultiGpuSearchIndex control plane on CPU: worker spawn, collective fan-out, shard merge, failure handling.

The per-GPU index is replaced by tests/fake_shard_index.py through ``index_factory``; the GPU path itself is covered by
tests/test_gpu_multirank.py::test_multigpu_serving (2 GPUs)."""
import numpy as np
import pytest

from infomesh_b200.engine.gpu_index import _slice_csr, format_hits, merge_shard_arrays
from infomesh_b200.engine.multigpu import MultiGpuSearchIndex, make_index
from infomesh_b200.index.local_store import LocalStore
from infomesh_b200.ops.bm25 import HostIndexBuilder

WORDS = ["tensor", "memory", "barrier", "switch", "kernel", "cluster", "pipeline", "shard", "vector", "search"]


def _store(tmp_path, n=23):
    st = LocalStore(tmp_path / "index.db")
    for i in range(n):
        body = " ".join(WORDS[(i + j) % len(WORDS)] for j in range(3 + i % 4)) + f" unique{i} " * (1 + i % 2)
        st.add_document(url=f"https://example.org/p/{i}", title=f"Page {i}", text=body, raw_html_hash=f"h{i}", text_hash=f"t{i}")
    return st


@pytest.fixture
def front(tmp_path):
    st = _store(tmp_path)
    f = MultiGpuSearchIndex(st, devices=2, store_path=str(tmp_path / "index.db"), query_batch=4,
                            index_factory="fake_shard_index:make", call_timeout=30, ready_timeout=60)
    yield f, st
    f.close()
    st.close()


def test_rebuild_counts_every_shard(front):
    f, st = front
    assert f.rebuild() == 23
    s = f.stats()
    assert s["gpus"] == 2 and s["healthy"] and s["documents"] == 23
    assert [p["documents"] for p in s["per_rank"]] == [12, 11]
    assert s["hbm_bytes"] == 2001


def test_search_merges_rows_owned_by_different_ranks(front):
    f, st = front
    f.rebuild()
    hits = f.search("unique3", k=5)
    assert [h["url"] for h in hits] == ["https://example.org/p/3"]          # owned by rank 0
    hits = f.search("unique20", k=5)
    assert [h["url"] for h in hits] == ["https://example.org/p/20"]         # owned by rank 1
    assert hits[0]["snippet"] == "unique20"                                 # span chosen by the owning rank
    many = f.search_many(["tensor memory", "unique1", "unique22", "nothingmatches", "kernel"], k=3)
    assert len(many) == 5 and many[3] == []
    assert {h["url"] for h in many[1]} == {"https://example.org/p/1"} and {h["url"] for h in many[2]} == {"https://example.org/p/22"}
    # both halves of the corpus show up in one ranked list
    urls = [int(h["url"].rsplit("/", 1)[1]) for q in ("tensor", "search", "kernel") for h in f.search(q, k=5)]
    assert min(urls) < 12 <= max(urls)


def test_delete_reaches_the_owning_shard_and_surface_attributes(front):
    f, st = front
    assert f.engine is None and not f.mark_deleted(1)          # not serving yet
    f.rebuild()
    assert f.engine is not None and f.reranker is None
    doc20 = f.search("unique20", k=1)[0]["doc_id"]
    assert f.mark_deleted(doc20) and not f.mark_deleted(10 ** 9)
    assert f.search("unique20", k=1) == []
    f.note_added(3)
    assert f.refresh() == 0                                    # rebuild in place: nothing new in the store


def test_worker_crash_marks_front_unhealthy_and_rebuild_recovers(tmp_path):
    st = _store(tmp_path)
    f = MultiGpuSearchIndex(st, devices=2, store_path=str(tmp_path / "index.db"), query_batch=4,
                            index_factory="fake_shard_index:make", call_timeout=10, ready_timeout=60, fail_on=1)
    try:
        f.rebuild()
        assert f.search("tensor")
        assert f.search("crash now") == []
        assert not f.healthy and f.search("tensor") == []
        assert f.rebuild() == 23 and f.healthy and f.search("tensor")
    finally:
        f.close()
        st.close()


def test_per_shard_segments_round_trip(front, tmp_path):
    f, st = front
    f.rebuild()
    before = f.search_many(["unique3", "unique20", "tensor kernel"], k=5)
    man = f.save(tmp_path / "segments")
    assert man["world"] == 2 and man["n_docs"] == 23 and [s["dir"] for s in man["shards"]] == ["shard-00-of-02", "shard-01-of-02"]
    assert (tmp_path / "segments" / "shard-01-of-02" / "fake.json").exists()
    # warm group: every worker swaps in its loaded shard
    assert f.load(tmp_path / "segments") == 23
    assert f.search_many(["unique3", "unique20", "tensor kernel"], k=5) == before
    # cold start: a fresh front spawns its workers straight from the segments (no rebuild)
    g = MultiGpuSearchIndex(st, devices=2, store_path=str(tmp_path / "index.db"), query_batch=4,
                            index_factory="fake_shard_index:make", call_timeout=30, ready_timeout=60)
    try:
        assert g.load(tmp_path / "segments") == 23 and g.healthy
        assert g.search_many(["unique3", "unique20", "tensor kernel"], k=5) == before
    finally:
        g.close()


def test_segments_for_another_gpu_count_are_refused(front, tmp_path):
    f, st = front
    f.rebuild()
    f.save(tmp_path / "segments")
    g = MultiGpuSearchIndex(st, devices=4, store_path=str(tmp_path / "index.db"), query_batch=4,
                            index_factory="fake_shard_index:make", call_timeout=30, ready_timeout=60)
    with pytest.raises(ValueError, match="written for 2 GPUs"):
        g.load(tmp_path / "segments")
    g.close()


def test_warm_start_loads_fresh_segments_and_rebuilds_stale_ones(front, tmp_path):
    from infomesh_b200.engine.multigpu import warm_start

    f, st = front
    seg = tmp_path / "seg"
    assert warm_start(f, st, seg) == 23 and (seg / "manifest.json").exists()          # first start: build + save
    calls = []
    real_rebuild, real_load = f.rebuild, f.load
    f.rebuild = lambda: (calls.append("rebuild"), real_rebuild())[1]
    f.load = lambda d: (calls.append("load"), real_load(d))[1]
    assert warm_start(f, st, seg) == 23 and calls == ["load"]                          # unchanged store: files only
    st.add_document(url="https://example.org/p/new", title="New", text="tensor unique99 body text", raw_html_hash="hn", text_hash="tn")
    assert warm_start(f, st, seg) == 24 and calls == ["load", "rebuild"]               # grown store: stale segments are not served
    assert [h["url"] for h in f.search("unique99", k=5)] == ["https://example.org/p/new"]
    assert warm_start(f, st, seg) == 24 and calls[-1] == "load"                        # and the rebuild refreshed the files
    assert warm_start(f, st, "") == 24 and calls[-1] == "rebuild"                      # no directory configured: plain build


def test_failed_worker_start_raises(tmp_path):
    st = _store(tmp_path, 4)
    f = MultiGpuSearchIndex(st, devices=2, store_path=str(tmp_path / "index.db"), query_batch=4,
                            index_factory="fake_shard_index:does_not_exist", ready_timeout=60)
    with pytest.raises(RuntimeError, match="GPU worker 0"):
        f.rebuild()
    assert not f.healthy
    st.close()


def test_constructor_validation(tmp_path):
    st = _store(tmp_path, 2)
    with pytest.raises(ValueError):
        MultiGpuSearchIndex(st, devices=1)
    with pytest.raises(ValueError):
        MultiGpuSearchIndex(st, devices=3, query_batch=64)
    st.close()


def test_merge_and_format():
    a = {"scores": np.array([[3.0, 2.0, 1.0]], np.float32), "rows": np.array([[0, 5, 2]]), "doc_ids": np.array([[10, -1, 12]]),
         "pass": np.array([[0, -1, -1]]), "span": np.array([[[0, 4], [-1, -1], [-1, -1]]])}
    b = {"scores": a["scores"], "rows": a["rows"], "doc_ids": np.array([[-1, 55, -1]]), "pass": np.array([[-1, 1, -1]]),
         "span": np.array([[[-1, -1], [2, 6], [-1, -1]]])}
    m = merge_shard_arrays([a, b])
    # the owning rank may hold the row in another slot (near-ties resolved differently): matched by row value
    b2 = {"scores": a["scores"], "rows": np.array([[5, 0, 2]]), "doc_ids": np.array([[55, -1, -1]]), "pass": np.array([[1, -1, -1]]),
          "span": np.array([[[2, 6], [-1, -1], [-1, -1]]])}
    m2 = merge_shard_arrays([a, b2])
    assert m2["doc_ids"].tolist() == [[10, 55, 12]] and m2["span"][0, 1].tolist() == [2, 6]
    assert m["doc_ids"].tolist() == [[10, 55, 12]] and m["pass"].tolist() == [[0, 1, -1]] and m["span"][0, 1].tolist() == [2, 6]

    class Doc:
        def __init__(self, i):
            self.doc_id, self.url, self.title, self.text, self.crawled_at = i, f"u{i}", f"t{i}", "abcdefghij needle tail", 1.0

    class S:
        def get_document(self, i):
            return None if i == 12 else Doc(i)

    hits = format_hits(S(), ["needle"], 10, m)[0]
    assert [h["doc_id"] for h in hits] == [10, 55] and hits[0]["snippet"] == "abcd" and hits[1]["snippet"] == "cdef"


def test_slice_csr_matches_a_direct_build_of_the_range():
    texts = [f"alpha beta w{i % 7} gamma{i % 3} " * (1 + i % 3) for i in range(40)]
    g = HostIndexBuilder()
    for t in texts:
        g.add_text(t)
    glob = g.export()
    loc = _slice_csr(glob, 10, 25)
    d = HostIndexBuilder()
    for t in texts:
        d.tokenize(t, add=True)
    for t in texts[10:25]:
        d.add_text(t)
    ref = d.export()
    for term in range(len(glob["off"]) - 1):
        sl, sr = slice(loc["off"][term], loc["off"][term + 1]), slice(ref["off"][term], ref["off"][term + 1])
        assert np.array_equal(loc["doc"][sl], ref["doc"][sr]) and np.array_equal(loc["tf"][sl], ref["tf"][sr])
    assert loc["n_docs_global"] == 40 and np.array_equal(loc["df_global"], glob["df"]) and np.array_equal(loc["doc_len"], ref["doc_len"])


def test_make_index_single_gpu_when_devices_is_one(tmp_path, monkeypatch):
    import infomesh_b200.engine.gpu_index as gi

    made = {}
    monkeypatch.setattr(gi, "GpuSearchIndex", lambda store, **kw: made.update(kw) or "single")

    class G:
        devices, device, query_batch, encoder_path, reranker_path, allow_untrained_models = 1, 3, 64, "", "", False

    st = _store(tmp_path, 2)
    assert make_index(st, G()) == "single" and made["device"] == "cuda:3"
    st.close()
