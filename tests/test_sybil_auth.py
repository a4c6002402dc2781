"""p2p/sybil.py (PoW, subnet quotas) and p2p/message_auth.py (signed envelopes, replay / freshness / isolation)."""
import dataclasses

import pytest

from infomesh_b200.p2p import message_auth as MA
from infomesh_b200.p2p import sybil as SY
from infomesh_b200.p2p.keys import KeyPair


def test_leading_zero_bits():
    assert SY.leading_zero_bits(b"\x00\x00\xff") == 16 and SY.leading_zero_bits(b"\x01") == 7
    assert SY.leading_zero_bits(b"\x00\x00") == 16 and SY.leading_zero_bits(b"\x80") == 0


def test_pow_generate_verify_and_node_id():
    pub = b"k" * 32
    pw = SY.generate_pow(pub, difficulty_bits=10)
    assert SY.verify_pow(pub, pw.nonce, 10) and pw.difficulty_bits == 10 and len(pw.hash_hex) == 64
    assert SY.compute_pow_hash(pub, pw.nonce).hex() == pw.hash_hex
    assert SY.derive_node_id(pub, pw.nonce) == pw.hash_hex[:40]
    assert not SY.verify_pow(b"x" * 32, pw.nonce, 30)
    with pytest.raises(RuntimeError):
        SY.generate_pow(pub, difficulty_bits=60, max_nonce=50)


def test_subnet_of_v4_and_v6():
    assert SY.subnet_of("10.1.2.3") == "10.1.2.0/24"
    assert SY.subnet_of("2001:db8:abcd:12::1") == "2001:db8:abcd::/48"
    with pytest.raises(ValueError):
        SY.subnet_of("not-an-ip")


def test_subnet_limiter_quota_per_bucket():
    lim = SY.SubnetLimiter(max_per_subnet=2)
    assert lim.add("10.0.0.1", "a", 5) and lim.add("10.0.0.2", "b", 5) and lim.add("10.0.0.1", "a", 5)   # idempotent
    assert not lim.can_add("10.0.0.9", 5) and not lim.add("10.0.0.9", "c", 5)
    assert lim.add("10.0.0.9", "c", 6) and lim.add("10.0.1.9", "d", 5)                                    # other bucket / subnet
    assert lim.get_subnet_counts(5) == {"10.0.0.0/24": 2, "10.0.1.0/24": 1} and lim.total_nodes() == 4
    lim.remove("10.0.0.1", "a", 5)
    assert lim.can_add("10.0.0.9", 5)
    lim.remove("10.0.0.9", "c", 6)
    assert lim.get_subnet_counts(6) == {}
    lim.remove("10.9.9.9", "zz", 42)                                                                      # unknown: no-op


def test_sybil_validator_orders_its_checks():
    pub = b"p" * 32
    pw = SY.generate_pow(pub, difficulty_bits=8)
    pid = SY.derive_node_id(pub, pw.nonce)
    v = SY.SybilValidator(difficulty_bits=8, max_per_subnet=1)
    assert v.validate_peer(pub, pw.nonce, "10.0.0.1", pid, 0) == (True, "ok")
    assert v.validate_peer(pub, pw.nonce, "10.0.0.1", "f" * 40, 0) == (False, "node_id_mismatch")
    bad_nonce = next(n for n in range(1000) if not SY.verify_pow(pub, n, 8))
    assert v.validate_peer(pub, bad_nonce, "10.0.0.1", pid, 0) == (False, "invalid_pow")
    pub2 = b"q" * 32
    pw2 = SY.generate_pow(pub2, difficulty_bits=8)
    assert v.validate_peer(pub2, pw2.nonce, "10.0.0.77", SY.derive_node_id(pub2, pw2.nonce), 0) == (False, "subnet_limit")


# ------------------------------------------------------------------ signed envelopes
def setup_pair():
    kp = KeyPair.generate()
    reg = MA.PeerKeyRegistry()
    reg.register(kp.peer_id, kp.public_key_bytes())
    return kp, reg


def test_sign_verify_roundtrip_and_dict_codec():
    kp, reg = setup_pair()
    ctr, trk = MA.NonceCounter(), MA.NonceTracker()
    env = MA.sign_envelope(b"hello", kp, ctr, now=1000.0)
    assert env.nonce == 1 and ctr.current == 1
    env2 = MA.envelope_from_dict(MA.envelope_to_dict(env))
    assert env2 == env and MA.verify_envelope(env2, reg, trk, now=1001.0) == b"hello"
    assert trk.highest(kp.peer_id) == 1 and kp.peer_id in reg and len(reg) == 1


def test_replay_and_stale_messages_rejected():
    kp, reg = setup_pair()
    ctr, trk = MA.NonceCounter(), MA.NonceTracker()
    e1 = MA.sign_envelope(b"a", kp, ctr, now=1000.0)
    e2 = MA.sign_envelope(b"b", kp, ctr, now=1000.0)
    assert MA.verify_envelope(e2, reg, trk, now=1000.0) == b"b"
    # frames of one sender may overtake each other (one TCP connection per request): an UNSEEN lower nonce inside the
    # sliding window is accepted once ...
    assert MA.verify_envelope(e1, reg, trk, now=1000.0) == b"a"
    # ... but every exact replay is refused, and so is anything older than the window
    for env in (e1, e2):
        with pytest.raises(MA.VerificationError, match="replayed"):
            MA.verify_envelope(env, reg, trk, now=1000.0)
    far = MA.NonceCounter(10 * MA.NONCE_WINDOW)
    ehi = MA.sign_envelope(b"hi", kp, far, now=1000.0)
    assert MA.verify_envelope(ehi, reg, trk, now=1000.0) == b"hi"
    stale = MA.sign_envelope(b"old", kp, MA.NonceCounter(5), now=1000.0)
    with pytest.raises(MA.VerificationError, match="replayed"):
        MA.verify_envelope(stale, reg, trk, now=1000.0)
    ctr = MA.NonceCounter(20 * MA.NONCE_WINDOW)
    e3 = MA.sign_envelope(b"c", kp, ctr, now=1000.0)
    with pytest.raises(MA.VerificationError, match="too old"):
        MA.verify_envelope(e3, reg, trk, now=1000.0 + 10_000)


def test_forged_signature_does_not_burn_the_nonce():
    kp, reg = setup_pair()
    ctr, trk = MA.NonceCounter(), MA.NonceTracker()
    good = MA.sign_envelope(b"pay", kp, ctr, now=50.0)
    forged = dataclasses.replace(good, payload=b"PAY", nonce=7)
    with pytest.raises(MA.VerificationError, match="invalid signature"):
        MA.verify_envelope(forged, reg, trk, now=50.0)
    assert trk.highest(kp.peer_id) == 0                                  # victim's range untouched
    assert MA.verify_envelope(good, reg, trk, now=50.0) == b"pay"


def test_unknown_and_isolated_peers():
    kp, reg = setup_pair()
    env = MA.sign_envelope(b"x", kp, MA.NonceCounter(), now=5.0)
    with pytest.raises(MA.VerificationError, match="unknown public key"):
        MA.verify_envelope(env, MA.PeerKeyRegistry(), MA.NonceTracker(), now=5.0)
    with pytest.raises(MA.VerificationError, match="isolated"):
        MA.verify_envelope(env, reg, MA.NonceTracker(), now=5.0, is_isolated_fn=lambda p: True)
    reg.remove(kp.peer_id)
    assert reg.get(kp.peer_id) is None


def test_nonce_tracker_evicts_least_recent_senders():
    trk = MA.NonceTracker()
    for i in range(MA.MAX_NONCE_HISTORY + 5):
        assert trk.check_and_record(f"peer{i}", 1)
    assert trk.highest("peer0") == 0 and trk.highest(f"peer{MA.MAX_NONCE_HISTORY + 4}") == 1
    assert not trk.check_and_record(f"peer{MA.MAX_NONCE_HISTORY + 4}", 1)
