"""Property-based tests (hypothesis) for the pure cores the rest of the system leans on.  The reference has none
("no property-based tests"); these pin invariants rather than examples."""
import string

from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from infomesh_b200.compression.zstd import Compressor
from infomesh_b200.crawler.dedup import normalize_url
from infomesh_b200.crawler.simhash import hamming_distance, is_near_duplicate, simhash
from infomesh_b200.p2p import protocol as P
from infomesh_b200.p2p.replication import replica_peers
from infomesh_b200.search.query import sanitize_fts_query
from infomesh_b200.trust.merkle import MerkleTree, deserialize_proof, serialize_proof

FAST = settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
words = st.text(alphabet=string.ascii_lowercase, min_size=1, max_size=9)
hexes = st.text(alphabet="0123456789abcdef", min_size=64, max_size=64)


@FAST
@given(st.lists(hexes, min_size=1, max_size=40, unique=True), st.data())
def test_every_merkle_leaf_has_a_proof_that_verifies_and_no_other_document_does(leaves, data):
    tree = MerkleTree()
    root = tree.build(leaves)
    i = data.draw(st.integers(0, len(leaves) - 1))
    proof = tree.get_proof(i)
    assert proof.root_hash == root and MerkleTree.verify_proof(proof) and MerkleTree.verify_document(leaves[i], proof)
    assert MerkleTree.verify_proof(deserialize_proof(serialize_proof(proof)))                 # survives the wire form
    other = data.draw(hexes)
    if other != leaves[i]:
        assert not MerkleTree.verify_document(other, proof)
    assert MerkleTree().build(leaves) == root and (len(leaves) == 1 or MerkleTree().build(leaves[::-1]) != root or leaves == leaves[::-1])


@FAST
@given(st.sampled_from(list(P.MessageType)), st.dictionaries(words, st.one_of(st.integers(-2 ** 40, 2 ** 40), st.floats(allow_nan=False, allow_infinity=False), st.text(max_size=40),
                                                                              st.binary(max_size=40), st.booleans(), st.none(), st.lists(st.integers(0, 9), max_size=5)),
                                                              max_size=8))
def test_protocol_frames_round_trip_and_announce_their_own_length(msg_type, payload):
    frame = P.encode_message(msg_type, payload)
    assert P.read_frame_length(frame[:4]) == len(frame) - 4
    kind, back = P.decode_message(frame)
    assert kind == msg_type and back == payload


@FAST
@given(st.lists(words, min_size=1, max_size=4), st.lists(st.tuples(words, words), max_size=4),
       st.sampled_from(["", "utm_source=x", "fbclid=1", "gclid=2&utm_medium=y"]), st.sampled_from(["", "#frag", "#a/b"]))
def test_url_normalisation_is_idempotent_and_blind_to_tracking_noise(path, query, tracking, fragment):
    q = "&".join(f"{k}={v}" for k, v in query)
    base = "https://Example.ORG/" + "/".join(path)
    noisy = base + "?" + "&".join(x for x in (q, tracking) if x) + fragment
    clean = base + ("?" + q if q else "")
    once = normalize_url(noisy)
    assert normalize_url(once) == once                                                       # a fixed point
    if not any(k in ("ref", "source") or k.startswith("utm_") or k in ("fbclid", "gclid", "mc_cid", "mc_eid") for k, _ in query):
        assert once == normalize_url(clean)                                                  # tracking params, fragment, host case: no effect
    assert "#" not in once and once.startswith("https://example.org/")


@FAST
@given(st.lists(words, min_size=12, max_size=60), st.data())
def test_simhash_is_deterministic_and_a_small_edit_moves_it_less_than_a_rewrite(doc, data):
    text = " ".join(doc)
    a = simhash(text)
    assert a == simhash(text) and 0 <= a < 2 ** 64 and hamming_distance(a, a) == 0 and is_near_duplicate(a, a)
    other = " ".join(data.draw(st.lists(words, min_size=12, max_size=60)))
    assert hamming_distance(a, simhash(other)) == hamming_distance(simhash(other), a) <= 64


@FAST
@given(st.binary(max_size=5000), st.integers(1, 19))
def test_zstd_frames_round_trip_at_every_level(blob, level):
    c = Compressor(level=level)
    assert c.decompress(c.compress(blob)) == blob and Compressor(level=3).decompress(c.compress(blob)) == blob


@FAST
@given(st.text(max_size=300))
def test_the_fts_sanitiser_never_returns_an_empty_or_operator_bearing_query(raw):
    out = sanitize_fts_query(raw)
    assert out and out == " ".join(out.split()) and not set(out) & set('"()*{}^:')
    assert sanitize_fts_query(out) == out or out == "infomesh"


@FAST
@given(st.text(min_size=1, max_size=60), st.lists(words, min_size=1, max_size=12, unique=True), st.integers(1, 6))
def test_replica_placement_is_stable_under_candidate_order_and_growth(url, peers, n):
    chosen = replica_peers(url, peers, n)
    assert len(chosen) == min(n, len(peers)) and len(set(chosen)) == len(chosen) and set(chosen) <= set(peers)
    assert replica_peers(url, peers[::-1], n) == chosen                                      # order of the peer list does not matter
    assert replica_peers(url, peers, n + 1)[:len(chosen)] == chosen                          # asking for more keeps the closer ones first


# ----------------------------------------------------------------------------- sharding invariants of the multi-GPU index (host side)
import numpy as np  # noqa: E402

from infomesh_b200.engine.gpu_index import _slice_csr, merge_shard_arrays  # noqa: E402


@FAST
@given(st.integers(1, 40), st.integers(1, 12), st.integers(1, 5), st.integers(0, 2 ** 31 - 1))
def test_posting_slices_partition_the_global_index(n_docs, vocab, world, seed):
    """Every posting of the global CSR lands in exactly one shard, re-based to that shard's rows, and every shard carries the
    GLOBAL statistics (so BM25 scores do not depend on the number of GPUs)."""
    rng = np.random.default_rng(seed)
    postings = [sorted(rng.choice(n_docs, size=int(rng.integers(0, n_docs + 1)), replace=False).tolist()) for _ in range(vocab)]
    off = np.zeros(vocab + 1, np.int64)
    off[1:] = np.cumsum([len(p) for p in postings])
    doc = np.array([d for p in postings for d in p], dtype=np.int32)
    tf = rng.integers(1, 9, size=doc.size).astype(np.uint8)
    csr = {"off": off, "doc": doc, "tf": tf, "doc_len": rng.integers(1, 50, size=n_docs).astype(np.int32), "df": np.diff(off).astype(np.int32)}
    per = (n_docs + world - 1) // world
    seen = [[] for _ in range(vocab)]
    for r in range(world):
        lo, hi = r * per, min(n_docs, (r + 1) * per)
        if lo >= hi:
            continue
        part = _slice_csr(csr, lo, hi)
        assert part["n_docs_global"] == n_docs and (part["df_global"] == csr["df"]).all() and len(part["doc_len"]) == hi - lo
        assert abs(part["avg_len_global"] - csr["doc_len"].mean()) < 1e-9
        for t in range(vocab):
            rows = part["doc"][part["off"][t]:part["off"][t + 1]]
            assert (rows >= 0).all() and (rows < hi - lo).all() and (np.diff(rows) > 0).all()          # local rows, still ascending
            seen[t] += [(int(x) + lo, int(f)) for x, f in zip(rows, part["tf"][part["off"][t]:part["off"][t + 1]])]
    for t in range(vocab):
        want = [(int(d), int(f)) for d, f in zip(doc[off[t]:off[t + 1]], tf[off[t]:off[t + 1]])]
        assert seen[t] == want


@FAST
@given(st.integers(1, 6), st.integers(1, 8), st.integers(2, 5), st.integers(0, 2 ** 31 - 1))
def test_shard_views_merge_to_the_complete_answer(nq, k, world, seed):
    """Rank 0 decides scores and rows; every other field comes from whichever rank owns the row -- wherever that rank happened
    to place the row among near-ties."""
    rng = np.random.default_rng(seed)
    n_rows = k * 4
    rows = np.stack([rng.choice(n_rows, size=k, replace=False) for _ in range(nq)]).astype(np.int64)
    rows[rng.random(rows.shape) < 0.15] = -1                                   # empty result slots
    owner = rng.integers(0, world, size=n_rows)
    scores = -np.sort(-rng.random((nq, k)).astype(np.float32), axis=1)
    parts = []
    for r in range(world):
        perm = np.stack([rng.permutation(k) for _ in range(nq)]) if r else np.tile(np.arange(k), (nq, 1))     # ranks > 0 order slots their own way
        view = np.take_along_axis(rows, perm, axis=1)
        mine = (view >= 0) & (owner[np.clip(view, 0, None)] == r)
        parts.append({"scores": np.take_along_axis(scores, perm, axis=1), "rows": view, "doc_ids": np.where(mine, 1000 + view, -1),
                      "pass": np.where(mine, view % 3, -1), "span": np.where(mine[..., None], np.stack([view, view + 5], axis=-1), -1)})
    merged = merge_shard_arrays(parts)
    assert (merged["rows"] == rows).all() and (merged["scores"] == scores).all()
    live = rows >= 0
    assert (merged["doc_ids"][live] == 1000 + rows[live]).all() and (merged["doc_ids"][~live] == -1).all()
    assert (merged["pass"][live] == rows[live] % 3).all() and (merged["span"][live][:, 1] == rows[live] + 5).all()
