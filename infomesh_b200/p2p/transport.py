"""Stream transport for the peer protocols.

The reference rides on py-libp2p (trio) streams (infomesh/p2p/node.py:535-598, 771-981).  Neither libp2p nor trio is
a dependency here: peers speak the same ``[u32 BE length][msgpack {type, payload}]`` frames
(infomesh/p2p/protocol.py:303-378) over asyncio TCP, one short-lived connection per request like a libp2p stream.
When the node has a key pair a connection starts with the ``IMN1`` handshake of :mod:`infomesh_b200.p2p.secure_channel`
(ephemeral X25519 + Ed25519 identities, ChaCha20-Poly1305 records -- the role Noise plays under libp2p,
infomesh/p2p/node.py:550-552): frames are encrypted and the channel itself authenticates the sender.  Peers that do
not speak it are still served in plaintext with every frame wrapped in a :class:`SignedEnvelope` (unless
``require_encrypted``).  Bandwidth is charged to the token buckets; handlers are registered per :class:`MessageType`.

Multiaddrs are the ``/ip4/<host>/tcp/<port>/p2p/<peer_id>`` subset.
"""
from __future__ import annotations

import asyncio
import time
from dataclasses import dataclass
from typing import Any, Awaitable, Callable

from infomesh_b200.p2p import message_auth as MA
from infomesh_b200.p2p import secure_channel as SC
from infomesh_b200.p2p.protocol import (MAX_MESSAGE_SIZE, MessageType, decode_message, encode_message,
                                        encode_signed_envelope, read_frame_length, safe_unpackb)
from infomesh_b200.p2p.throttle import BandwidthThrottle
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)
SESSION_IDLE_S = 60.0          # an encrypted session without traffic is closed by the listener after this long

Handler = Callable[[dict[str, Any], "PeerInfo"], Awaitable[tuple[MessageType, dict[str, Any]] | None]]


@dataclass(frozen=True)
class PeerInfo:
    peer_id: str
    host: str
    port: int

    @property
    def multiaddr(self) -> str:
        return format_multiaddr(self.host, self.port, self.peer_id)


def format_multiaddr(host: str, port: int, peer_id: str = "") -> str:
    fam = "ip6" if ":" in host else "ip4"
    base = f"/{fam}/{host}/tcp/{port}"
    return f"{base}/p2p/{peer_id}" if peer_id else base


def parse_multiaddr(addr: str) -> tuple[str, int, str]:
    """-> (host, port, peer_id); also accepts ``host:port``."""
    if addr.startswith("/"):
        parts = addr.strip("/").split("/")
        kv = dict(zip(parts[::2], parts[1::2]))
        host = kv.get("ip4") or kv.get("ip6") or kv.get("dns4") or kv.get("dns") or ""
        if not host or "tcp" not in kv:
            raise ValueError(f"unsupported multiaddr: {addr}")
        return host, int(kv["tcp"]), kv.get("p2p", "")
    host, _, port = addr.rpartition(":")
    if not host or not port.isdigit():
        raise ValueError(f"unsupported address: {addr}")
    return host, int(port), ""


async def read_frame(reader: asyncio.StreamReader, *, max_bytes: int = MAX_MESSAGE_SIZE) -> bytes:
    prefix = await reader.readexactly(4)
    n = read_frame_length(prefix)
    if n > max_bytes:
        raise ValueError(f"frame too large: {n}")
    return prefix + await reader.readexactly(n)


# message types that change a node's state: an unsigned frame of these kinds is refused even when the transport otherwise
# accepts unsigned traffic (read-only requests from anonymous peers)
WRITE_TYPES = frozenset({MessageType.INDEX_SUBMIT, MessageType.REPLICATE_REQUEST, MessageType.DHT_STORE, MessageType.INDEX_PUBLISH,
                         MessageType.KEY_REVOCATION, MessageType.ATTESTATION_PUBLISH})


class Transport:
    def __init__(self, key_pair: Any | None = None, *, throttle: BandwidthThrottle | None = None,
                 is_isolated_fn: Callable[[str], bool] | None = None, require_signed: bool = False, encrypt: bool = True,
                 require_encrypted: bool = False):
        self.key_pair = key_pair
        self.encrypt = bool(encrypt and key_pair is not None)          # offer / accept the IMN1 encrypted channel
        self.require_encrypted = bool(require_encrypted)               # refuse plaintext peers altogether
        self._plain_peers: set[tuple[str, int]] = set()                # peers that turned the handshake down
        self._links: dict[tuple[str, int], "_Link"] = {}               # idle encrypted sessions, one per peer address
        self._link_locks: dict[tuple[str, int], asyncio.Lock] = {}
        self.handshakes_out = 0
        self._inbound: set[asyncio.StreamWriter] = set()               # accepted connections still open
        self.encrypted_in = self.encrypted_out = 0
        self.peer_id = key_pair.peer_id if key_pair is not None else ""
        self.throttle = throttle
        self.keys = MA.PeerKeyRegistry()
        self._nonces_in = MA.NonceTracker()
        self._nonce_out = MA.NonceCounter(int(time.time() * 1000))   # monotonic across restarts
        self._isolated = is_isolated_fn
        self._require_signed = require_signed
        self._handlers: dict[MessageType, Handler] = {}
        self._server: asyncio.AbstractServer | None = None
        self.host, self.port = "127.0.0.1", 0
        self.bytes_in = self.bytes_out = 0

    # ------------------------------------------------------------------ server side
    def register(self, msg_type: MessageType, handler: Handler) -> None:
        self._handlers[msg_type] = handler

    @property
    def registered(self) -> list[MessageType]:
        return sorted(self._handlers)

    async def listen(self, host: str = "0.0.0.0", port: int = 0) -> int:
        self._server = await asyncio.start_server(self._serve_conn, host, port)
        sock = self._server.sockets[0].getsockname()
        self.host, self.port = (host if host not in ("0.0.0.0", "::") else "127.0.0.1"), sock[1]
        return self.port

    def _drop_links(self) -> None:
        for link in self._links.values():
            link.close()
        self._links.clear()

    async def close(self) -> None:
        self._drop_links()
        for w in list(self._inbound):          # sessions that peers keep open would hold wait_closed() until they idle out
            w.close()
        if self._server is not None:
            self._server.close()
            await self._server.wait_closed()
            self._server = None

    def _wrap(self, frame: bytes) -> bytes:
        if self.key_pair is None:
            return frame
        env = MA.sign_envelope(frame, self.key_pair, self._nonce_out)
        d = MA.envelope_to_dict(env)
        d["public_key"] = self.key_pair.public_key_bytes()
        return encode_signed_envelope(d)

    def _unwrap(self, frame: bytes, channel_sender: str = "") -> tuple[MessageType, dict[str, Any], str]:
        """-> (type, payload, verified sender id or "").  ``channel_sender``: the peer id the encrypted channel already
        authenticated; a bare frame inside such a channel counts as signed by that peer."""
        kind, payload = decode_message(frame)
        if kind != MessageType.SIGNED_ENVELOPE:
            if channel_sender:
                return kind, payload, channel_sender
            if self._require_signed or kind in WRITE_TYPES:
                raise MA.VerificationError("unsigned message refused")
            return kind, payload, ""
        pid, pub = str(payload.get("peer_id", "")), payload.get("public_key")
        registry = self.keys
        first_use = isinstance(pub, bytes | bytearray) and bool(pid) and pid not in self.keys
        if first_use:
            from infomesh_b200.p2p.keys import peer_id_from_public_key

            if peer_id_from_public_key(bytes(pub)) != pid:      # trust-on-first-use, bound to the id derivation
                raise MA.VerificationError("public key does not match peer id")
            registry = MA.PeerKeyRegistry(1)                     # candidate key: verify FIRST, remember only on success
            registry.register(pid, bytes(pub))
        inner = MA.verify_envelope(MA.envelope_from_dict(payload), registry, self._nonces_in,
                                   is_isolated_fn=self._isolated)
        if first_use:
            self.keys.register(pid, bytes(pub))
        kind, body = decode_message(inner)
        return kind, body, pid

    async def _serve_conn(self, reader: asyncio.StreamReader, writer: asyncio.StreamWriter) -> None:
        peer = writer.get_extra_info("peername") or ("?", 0)
        session: SC.SecureSession | None = None
        self._inbound.add(writer)
        try:
            prefix = await asyncio.wait_for(reader.readexactly(4), timeout=30.0)
            if prefix == SC.MAGIC:
                if not self.encrypt:
                    return                                                  # no identity to answer with: drop the connection
                session = await SC.server_handshake(reader, writer, self.key_pair)
                if self._isolated is not None and self._isolated(session.remote_peer_id):
                    return
                self.keys.register(session.remote_peer_id, session.remote_public_key)      # proven by the handshake
                frame = await asyncio.wait_for(session.recv(reader), timeout=30.0)
                self.encrypted_in += 1
            else:
                if self.require_encrypted:
                    await self._send(writer, encode_message(MessageType.ERROR, {"error": "encrypted channel required"}))
                    return
                n = read_frame_length(prefix)
                if n > MAX_MESSAGE_SIZE:
                    raise ValueError(f"frame too large: {n}")
                frame = prefix + await asyncio.wait_for(reader.readexactly(n), timeout=30.0)
            async def answer(data: bytes) -> None:
                if session is not None:
                    if self.throttle:
                        await self.throttle.acquire_upload(len(data))
                    self.bytes_out += await session.send(writer, data)
                else:
                    await self._send(writer, data)

            while True:
                self.bytes_in += len(frame)
                if self.throttle:
                    await self.throttle.acquire_download(len(frame))
                try:
                    kind, payload, sender = self._unwrap(frame, channel_sender=session.remote_peer_id if session else "")
                except (MA.VerificationError, ValueError) as exc:
                    await answer(encode_message(MessageType.ERROR, {"error": str(exc)}))
                    return
                handler = self._handlers.get(kind)
                if handler is None:
                    await answer(encode_message(MessageType.ERROR, {"error": f"unsupported type {int(kind)}"}))
                    return
                reply = await handler(payload, PeerInfo(sender, peer[0], peer[1]))
                if reply is not None:
                    # inside an authenticated channel the reply needs no envelope of its own
                    await answer(encode_message(*reply) if session is not None else self._wrap(encode_message(*reply)))
                if session is None:
                    return                      # legacy plaintext peers: one request per connection, as before
                # an encrypted session stays open: the handshake (2 x X25519 + 2 x Ed25519) is paid once per peer, not per
                # request; the client closes it, or it is dropped after SESSION_IDLE_S without traffic
                try:
                    frame = await asyncio.wait_for(session.recv(reader), timeout=SESSION_IDLE_S)
                except (asyncio.TimeoutError, asyncio.IncompleteReadError, ConnectionError):
                    return
                self.encrypted_in += 1
                if self._isolated is not None and self._isolated(session.remote_peer_id):
                    return
        except (asyncio.IncompleteReadError, asyncio.TimeoutError, ConnectionError, ValueError, SC.HandshakeError) as exc:
            logger.debug("transport_conn_error", error=str(exc))
        except Exception:  # noqa: BLE001 — a handler bug must not kill the listener
            logger.exception("transport_handler_failed")
        finally:
            self._inbound.discard(writer)
            writer.close()
            try:
                await writer.wait_closed()
            except Exception:  # noqa: BLE001
                pass

    async def _send(self, writer: asyncio.StreamWriter, data: bytes) -> None:
        if self.throttle:
            await self.throttle.acquire_upload(len(data))
        writer.write(data)
        await writer.drain()
        self.bytes_out += len(data)

    # ------------------------------------------------------------------ client side
    async def request(self, addr: str | tuple[str, int], msg_type: MessageType, payload: dict[str, Any], *,
                      timeout: float = 5.0, expect_reply: bool = True,
                      expect_peer: str | None = None) -> tuple[MessageType, dict[str, Any]] | None:
        """One request / reply exchange.  ``expect_peer``: the peer id that was dialled -- a reply signed by anyone else
        (an on-path host answering with its own self-signed envelope) or not signed at all is rejected."""
        host, port = addr if isinstance(addr, tuple) else parse_multiaddr(addr)[:2]
        if expect_peer is None and not isinstance(addr, tuple):
            expect_peer = parse_multiaddr(addr)[2] or None             # /p2p/<peer_id> in the multiaddr pins the identity

        async def _dial():
            reader, writer = await asyncio.open_connection(host, port)
            try:
                session = await SC.client_handshake(reader, writer, self.key_pair, expect_peer=expect_peer)
            except BaseException:
                writer.close()
                raise
            self.keys.register(session.remote_peer_id, session.remote_public_key)
            self.handshakes_out += 1
            return _Link(reader, writer, session)

        async def _exchange(link: "_Link"):
            data = encode_message(msg_type, payload)
            if self.throttle:
                await self.throttle.acquire_upload(len(data))
            self.bytes_out += await link.session.send(link.writer, data)
            self.encrypted_out += 1
            if not expect_reply:
                return None
            frame = await link.session.recv(link.reader)
            self.bytes_in += len(frame)
            if self.throttle:
                await self.throttle.acquire_download(len(frame))
            kind, body, _sender = self._unwrap(frame, channel_sender=link.session.remote_peer_id)
            link.last_used = time.monotonic()
            return kind, body

        async def _secure():
            if not expect_reply:                  # fire-and-forget: a private connection, so a late reply cannot desynchronise
                link = await _dial()              # a session that other requests share
                try:
                    return await _exchange(link)
                finally:
                    link.close()
            key = (host, port)
            lock = self._link_locks.setdefault(key, asyncio.Lock())
            async with lock:                      # one request in flight per session: records are strictly ordered
                link = self._links.pop(key, None)
                if link is not None and (link.stale() or (expect_peer and link.session.remote_peer_id != expect_peer)):
                    link.close()
                    link = None
                reused = link is not None
                for attempt in (0, 1):
                    if link is None:
                        link = await _dial()
                    try:
                        out = await _exchange(link)
                    except (asyncio.IncompleteReadError, ConnectionError, SC.HandshakeError):
                        link.close()
                        link = None
                        if reused and attempt == 0:          # the peer dropped an idle session: dial again, once
                            reused = False
                            continue
                        raise
                    except BaseException:                     # timeout / cancellation mid-exchange: the stream position is unknown
                        link.close()
                        raise
                    self._links[key] = link
                    return out

        async def _plain():
            reader, writer = await asyncio.open_connection(host, port)
            try:
                await self._send(writer, self._wrap(encode_message(msg_type, payload)))
                if not expect_reply:
                    return None
                frame = await read_frame(reader)
                self.bytes_in += len(frame)
                if self.throttle:
                    await self.throttle.acquire_download(len(frame))
                kind, body, sender = self._unwrap(frame)
                if expect_peer and sender != expect_peer:
                    raise MA.VerificationError(f"reply signed by {sender[:16] or 'nobody'}, expected {expect_peer[:16]}")
                return kind, body
            finally:
                writer.close()

        async def _go():
            if self.encrypt and (host, port) not in self._plain_peers:
                try:
                    return await _secure()
                except (asyncio.IncompleteReadError, ConnectionError, SC.HandshakeError) as exc:
                    # an identity mismatch is final; a peer that merely does not speak IMN1 (it closes the connection or
                    # answers with a plaintext ERROR frame) is remembered and served the legacy way
                    if isinstance(exc, SC.HandshakeError) and "expected" in str(exc):
                        raise MA.VerificationError(str(exc)) from exc
                    if self.require_encrypted:
                        raise
                    self._plain_peers.add((host, port))
            elif self.require_encrypted:
                raise ConnectionError("peer does not support the encrypted channel")
            return await _plain()

        return await asyncio.wait_for(_go(), timeout=timeout)


class _Link:
    """An established outbound encrypted session, kept for the next request to the same peer."""

    def __init__(self, reader: asyncio.StreamReader, writer: asyncio.StreamWriter, session: "SC.SecureSession"):
        self.reader, self.writer, self.session = reader, writer, session
        self.last_used = time.monotonic()

    def stale(self) -> bool:
        return self.writer.is_closing() or time.monotonic() - self.last_used > SESSION_IDLE_S * 0.8 or self.reader.at_eof()

    def close(self) -> None:
        try:
            self.writer.close()
        except Exception:  # noqa: BLE001
            pass


__all__ = ["Transport", "PeerInfo", "format_multiaddr", "parse_multiaddr", "read_frame", "safe_unpackb"]
