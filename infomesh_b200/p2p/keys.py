"""Ed25519 node identity: generate / load / save (PEM, 0600), sign / verify, ``peer_id = sha256(pubkey)[:40]``,
key rotation with a dual-signed revocation record (reference infomesh/p2p/keys.py:41-416)."""
from __future__ import annotations

import contextlib
import hashlib
import os
import shutil
import stat
import time
from pathlib import Path

import msgpack
from cryptography.exceptions import InvalidSignature
from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PrivateKey, Ed25519PublicKey
from cryptography.hazmat.primitives.serialization import (Encoding, NoEncryption, PrivateFormat, PublicFormat,
                                                          load_pem_private_key)

from infomesh_b200.p2p.protocol import KeyRevocationRecord
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


class KeyPair:
    def __init__(self, private_key: Ed25519PrivateKey, public_key: Ed25519PublicKey):
        self._priv = private_key
        self._pub = public_key

    @classmethod
    def generate(cls) -> "KeyPair":
        priv = Ed25519PrivateKey.generate()
        return cls(priv, priv.public_key())

    @classmethod
    def load(cls, keys_dir: Path) -> "KeyPair":
        path = Path(keys_dir) / "private.pem"
        if not path.exists():
            raise FileNotFoundError(f"Private key not found: {path}")
        priv = load_pem_private_key(path.read_bytes(), password=None)
        if not isinstance(priv, Ed25519PrivateKey):
            raise ValueError("private.pem is not an Ed25519 key")
        return cls(priv, priv.public_key())

    def save(self, keys_dir: Path) -> None:
        d = Path(keys_dir)
        d.mkdir(parents=True, exist_ok=True)
        priv, pub = d / "private.pem", d / "public.pem"
        priv.write_bytes(self._priv.private_bytes(Encoding.PEM, PrivateFormat.PKCS8, NoEncryption()))
        os.chmod(priv, stat.S_IRUSR | stat.S_IWUSR)
        pub.write_bytes(self._pub.public_bytes(Encoding.PEM, PublicFormat.SubjectPublicKeyInfo))
        os.chmod(pub, stat.S_IRUSR | stat.S_IWUSR | stat.S_IRGRP | stat.S_IROTH)

    def sign(self, data: bytes) -> bytes:
        return self._priv.sign(data)

    def verify(self, data: bytes, signature: bytes) -> bool:
        try:
            self._pub.verify(signature, data)
            return True
        except (InvalidSignature, ValueError):
            return False

    def public_key_bytes(self) -> bytes:
        return self._pub.public_bytes(Encoding.Raw, PublicFormat.Raw)

    @property
    def peer_id(self) -> str:
        return hashlib.sha256(self.public_key_bytes()).hexdigest()[:40]


def verify_with_public_key(public_key: bytes, data: bytes, signature: bytes) -> bool:
    try:
        Ed25519PublicKey.from_public_bytes(public_key).verify(signature, data)
        return True
    except Exception:  # noqa: BLE001 — malformed key / signature are both "invalid"
        return False


def peer_id_from_public_key(public_key: bytes) -> str:
    return hashlib.sha256(public_key).hexdigest()[:40]


def ensure_keys(data_dir: Path) -> KeyPair:
    keys_dir = Path(data_dir) / "keys"
    nested = keys_dir / "keys"                    # early builds passed <data_dir>/keys here; pull that pair up one level
    if not (keys_dir / "private.pem").exists() and (nested / "private.pem").exists():
        for f in nested.iterdir():
            f.rename(keys_dir / f.name)
        with contextlib.suppress(OSError):
            nested.rmdir()
    if (keys_dir / "private.pem").exists():
        return KeyPair.load(keys_dir)
    logger.info("first_run_keygen", keys_dir=str(keys_dir))
    pair = KeyPair.generate()
    pair.save(keys_dir)
    return pair


def export_public_key(data_dir: Path) -> str:
    path = Path(data_dir) / "keys" / "public.pem"
    if not path.exists():
        raise FileNotFoundError(f"No public key found at {path}. Run 'infomesh start' first.")
    return path.read_text()


def _revocation_payload(r: KeyRevocationRecord) -> bytes:
    return msgpack.packb({"old_peer_id": r.old_peer_id, "new_peer_id": r.new_peer_id,
                          "old_public_key": r.old_public_key, "new_public_key": r.new_public_key, "reason": r.reason,
                          "timestamp": r.timestamp}, use_bin_type=True)


def rotate_keys(data_dir: Path, *, reason: str = "rotation") -> tuple[KeyPair, KeyPair, KeyRevocationRecord]:
    """Back the old pair up, mint a new one, and persist a revocation signed by BOTH keys."""
    keys_dir = Path(data_dir) / "keys"
    if not (keys_dir / "private.pem").exists():
        raise FileNotFoundError("No existing key pair to rotate. Run 'infomesh start' first.")
    old = KeyPair.load(keys_dir)
    stamp = time.time_ns()
    backup = None
    for n in range(1000):
        cand = keys_dir / (f"backup-{stamp}" + (f"-{n}" if n else ""))
        try:
            cand.mkdir(parents=True, exist_ok=False)
            backup = cand
            break
        except FileExistsError:
            continue
    if backup is None:
        raise FileExistsError("Could not create a unique key backup directory")
    for name in ("private.pem", "public.pem"):
        if (keys_dir / name).exists():
            shutil.copy2(keys_dir / name, backup / name)
    new = KeyPair.generate()
    draft = KeyRevocationRecord(old.peer_id, new.peer_id, old.public_key_bytes(), new.public_key_bytes(), reason,
                                time.time())
    payload = _revocation_payload(draft)
    record = KeyRevocationRecord(draft.old_peer_id, draft.new_peer_id, draft.old_public_key, draft.new_public_key,
                                 draft.reason, draft.timestamp, old.sign(payload), new.sign(payload))
    new.save(keys_dir)
    rev_dir = keys_dir / "revocations"
    rev_dir.mkdir(parents=True, exist_ok=True)
    (rev_dir / f"{old.peer_id[:16]}.bin").write_bytes(msgpack.packb(
        {**msgpack.unpackb(payload, raw=False), "old_key_signature": record.old_key_signature,
         "new_key_signature": record.new_key_signature}, use_bin_type=True))
    logger.info("keys_rotated", old_peer_id=old.peer_id, new_peer_id=new.peer_id)
    return old, new, record


def verify_revocation(record: KeyRevocationRecord) -> bool:
    payload = _revocation_payload(record)
    if peer_id_from_public_key(record.old_public_key) != record.old_peer_id:
        return False
    if peer_id_from_public_key(record.new_public_key) != record.new_peer_id:
        return False
    return (verify_with_public_key(record.old_public_key, payload, record.old_key_signature)
            and verify_with_public_key(record.new_public_key, payload, record.new_key_signature))


def load_revocations(data_dir: Path) -> list[KeyRevocationRecord]:
    """Every revocation record saved under ``<data_dir>/keys/revocations`` (oldest file name first); unreadable files are
    skipped (reference infomesh/p2p/keys.py:391-416)."""
    rev_dir = Path(data_dir) / "keys" / "revocations"
    if not rev_dir.is_dir():
        return []
    out: list[KeyRevocationRecord] = []
    for path in sorted(rev_dir.glob("*.bin")):
        try:
            raw = msgpack.unpackb(path.read_bytes(), raw=False)
            out.append(KeyRevocationRecord(**raw))
        except Exception:  # noqa: BLE001
            logger.warning("revocation_load_failed", path=str(path))
    return out
