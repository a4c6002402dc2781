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
