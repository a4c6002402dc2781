"""Encrypted, mutually authenticated sessions for the peer transport.

The reference gets confidentiality from libp2p's Noise security transport (infomesh/p2p/node.py:550-552); libp2p is not a
dependency here, so the TCP transport carries its own handshake built from the primitives in ``cryptography``:

* **Key agreement**: ephemeral X25519 on both sides (forward secrecy), HKDF-SHA256 over the shared secret with the
  handshake transcript as salt, one ChaCha20-Poly1305 key per direction.
* **Authentication** (SIGMA-I shape, what Noise XX + libp2p's signed identity payload gives): each side proves its
  long-term Ed25519 identity by signing the transcript hash; the identity blocks travel encrypted, so a passive observer
  learns neither peer id.  ``peer_id = sha256(ed25519_pub)[:40]`` as everywhere else in the mesh.
* **Framing**: after the three handshake messages every application frame is ``[u32 BE length][ciphertext || tag]`` with
  a 64-bit per-direction counter as nonce (a replayed, reordered or truncated record fails authentication).

Wire (client -> server first):

    C: "IMN1" || e_c(32)
    S: e_s(32) || AEAD(k_hs_s, identity_s)            identity = ed25519_pub(32) || sig(64) over H("IMN1" || e_c || e_s || role)
    C: AEAD(k_hs_c, identity_c)  then encrypted records ...

The magic ``IMN1`` read as a legacy length prefix would be a 1.2 GB frame (far above the 10 MiB cap), so a listener can
tell the two apart from the first four bytes and keep serving plaintext peers unless ``require_encrypted`` is set."""
from __future__ import annotations

import asyncio
import hashlib
import struct

from cryptography.hazmat.primitives import hashes
from cryptography.hazmat.primitives.asymmetric.x25519 import X25519PrivateKey, X25519PublicKey
from cryptography.hazmat.primitives.ciphers.aead import ChaCha20Poly1305
from cryptography.hazmat.primitives.kdf.hkdf import HKDF
from cryptography.hazmat.primitives.serialization import Encoding, PublicFormat

from infomesh_b200.p2p.keys import peer_id_from_public_key, verify_with_public_key

MAGIC = b"IMN1"
MAX_RECORD = 10 * 1024 * 1024 + 64
_ID_LEN = 32 + 64


class HandshakeError(Exception):
    """The peer failed to prove its identity, spoke another protocol, or the record stream was tampered with."""


def _kdf(shared: bytes, transcript: bytes, label: bytes) -> bytes:
    return HKDF(algorithm=hashes.SHA256(), length=32, salt=transcript, info=b"infomesh-b200 " + label).derive(shared)


def _nonce(counter: int) -> bytes:
    return b"\x00\x00\x00\x00" + struct.pack("<Q", counter)


class SecureSession:
    """Record layer of an established session (one instance per side)."""

    def __init__(self, send_key: bytes, recv_key: bytes, remote_peer_id: str, remote_public_key: bytes):
        self._tx, self._rx = ChaCha20Poly1305(send_key), ChaCha20Poly1305(recv_key)
        self._n_tx = self._n_rx = 0
        self.remote_peer_id, self.remote_public_key = remote_peer_id, remote_public_key

    def seal(self, plaintext: bytes) -> bytes:
        ct = self._tx.encrypt(_nonce(self._n_tx), plaintext, None)
        self._n_tx += 1
        return struct.pack(">I", len(ct)) + ct

    def open(self, ciphertext: bytes) -> bytes:
        try:
            pt = self._rx.decrypt(_nonce(self._n_rx), ciphertext, None)
        except Exception as exc:  # noqa: BLE001 -- InvalidTag
            raise HandshakeError("record failed authentication") from exc
        self._n_rx += 1
        return pt

    async def send(self, writer: asyncio.StreamWriter, plaintext: bytes) -> int:
        rec = self.seal(plaintext)
        writer.write(rec)
        await writer.drain()
        return len(rec)

    async def recv(self, reader: asyncio.StreamReader) -> bytes:
        (n,) = struct.unpack(">I", await reader.readexactly(4))
        if n > MAX_RECORD or n < 16:
            raise HandshakeError(f"bad record length {n}")
        return self.open(await reader.readexactly(n))


def _identity(key_pair, transcript_hash: bytes, role: bytes) -> bytes:
    return key_pair.public_key_bytes() + key_pair.sign(transcript_hash + role)


def _check_identity(blob: bytes, transcript_hash: bytes, role: bytes, expect_peer: str | None) -> tuple[str, bytes]:
    if len(blob) != _ID_LEN:
        raise HandshakeError("malformed identity block")
    pub, sig = blob[:32], blob[32:]
    if not verify_with_public_key(pub, transcript_hash + role, sig):
        raise HandshakeError("identity signature does not verify")
    pid = peer_id_from_public_key(pub)
    if expect_peer and pid != expect_peer:
        raise HandshakeError(f"connected to {pid[:16]}, expected {expect_peer[:16]}")
    return pid, pub


async def client_handshake(reader: asyncio.StreamReader, writer: asyncio.StreamWriter, key_pair, *,
                           expect_peer: str | None = None, timeout: float = 10.0) -> SecureSession:
    eph = X25519PrivateKey.generate()
    e_c = eph.public_key().public_bytes(Encoding.Raw, PublicFormat.Raw)
    writer.write(MAGIC + e_c)
    await writer.drain()
    head = await asyncio.wait_for(reader.readexactly(32 + _ID_LEN + 16), timeout)
    e_s, sealed = head[:32], head[32:]
    th = hashlib.sha256(MAGIC + e_c + e_s).digest()
    shared = eph.exchange(X25519PublicKey.from_public_bytes(e_s))
    k_hs_s, k_hs_c = _kdf(shared, th, b"hs s->c"), _kdf(shared, th, b"hs c->s")
    try:
        server_id = ChaCha20Poly1305(k_hs_s).decrypt(_nonce(0), sealed, th)
    except Exception as exc:  # noqa: BLE001
        raise HandshakeError("server identity block failed authentication") from exc
    pid, pub = _check_identity(server_id, th, b"server", expect_peer)
    writer.write(ChaCha20Poly1305(k_hs_c).encrypt(_nonce(0), _identity(key_pair, th, b"client"), th))
    await writer.drain()
    return SecureSession(_kdf(shared, th, b"app c->s"), _kdf(shared, th, b"app s->c"), pid, pub)


async def server_handshake(reader: asyncio.StreamReader, writer: asyncio.StreamWriter, key_pair, *, timeout: float = 10.0,
                           magic_consumed: bool = True) -> SecureSession:
    """Run the responder side.  ``magic_consumed``: the caller already read the 4 magic bytes to pick the protocol."""
    if not magic_consumed and await asyncio.wait_for(reader.readexactly(4), timeout) != MAGIC:
        raise HandshakeError("not an IMN1 handshake")
    e_c = await asyncio.wait_for(reader.readexactly(32), timeout)
    eph = X25519PrivateKey.generate()
    e_s = eph.public_key().public_bytes(Encoding.Raw, PublicFormat.Raw)
    th = hashlib.sha256(MAGIC + e_c + e_s).digest()
    shared = eph.exchange(X25519PublicKey.from_public_bytes(e_c))
    k_hs_s, k_hs_c = _kdf(shared, th, b"hs s->c"), _kdf(shared, th, b"hs c->s")
    writer.write(e_s + ChaCha20Poly1305(k_hs_s).encrypt(_nonce(0), _identity(key_pair, th, b"server"), th))
    await writer.drain()
    sealed = await asyncio.wait_for(reader.readexactly(_ID_LEN + 16), timeout)
    try:
        client_id = ChaCha20Poly1305(k_hs_c).decrypt(_nonce(0), sealed, th)
    except Exception as exc:  # noqa: BLE001
        raise HandshakeError("client identity block failed authentication") from exc
    pid, pub = _check_identity(client_id, th, b"client", None)
    return SecureSession(_kdf(shared, th, b"app s->c"), _kdf(shared, th, b"app c->s"), pid, pub)
