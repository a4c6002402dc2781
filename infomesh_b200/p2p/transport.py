"""Stream transport for the peer protocols.

The reference rides on py-libp2p (trio) streams (infomesh/p2p/node.py:535-598, 771-981).  Neither libp2p nor trio is
a dependency here: peers speak the same ``[u32 BE length][msgpack {type, payload}]`` frames
(infomesh/p2p/protocol.py:303-378) over plain asyncio TCP, one short-lived connection per request like a libp2p
stream.  Every request frame is wrapped in a :class:`SignedEnvelope` when the node has a key pair, bandwidth is
charged to the token buckets, and handlers are registered per :class:`MessageType`.

Multiaddrs are the ``/ip4/<host>/tcp/<port>/p2p/<peer_id>`` subset.
"""
from __future__ import annotations

import asyncio
import time
from dataclasses import dataclass
from typing import Any, Awaitable, Callable

from infomesh_b200.p2p import message_auth as MA
from infomesh_b200.p2p.protocol import (MAX_MESSAGE_SIZE, MessageType, decode_message, encode_message,
                                        encode_signed_envelope, read_frame_length, safe_unpackb)
from infomesh_b200.p2p.throttle import BandwidthThrottle
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

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
                 is_isolated_fn: Callable[[str], bool] | None = None, require_signed: bool = False):
        self.key_pair = key_pair
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

    async def close(self) -> None:
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

    def _unwrap(self, frame: bytes) -> tuple[MessageType, dict[str, Any], str]:
        """-> (type, payload, verified sender id or "")."""
        kind, payload = decode_message(frame)
        if kind != MessageType.SIGNED_ENVELOPE:
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
        try:
            frame = await asyncio.wait_for(read_frame(reader), timeout=30.0)
            self.bytes_in += len(frame)
            if self.throttle:
                await self.throttle.acquire_download(len(frame))
            try:
                kind, payload, sender = self._unwrap(frame)
            except (MA.VerificationError, ValueError) as exc:
                await self._send(writer, encode_message(MessageType.ERROR, {"error": str(exc)}))
                return
            handler = self._handlers.get(kind)
            if handler is None:
                await self._send(writer, encode_message(MessageType.ERROR, {"error": f"unsupported type {int(kind)}"}))
                return
            reply = await handler(payload, PeerInfo(sender, peer[0], peer[1]))
            if reply is not None:
                await self._send(writer, self._wrap(encode_message(*reply)))
        except (asyncio.IncompleteReadError, asyncio.TimeoutError, ConnectionError, ValueError) as exc:
            logger.debug("transport_conn_error", error=str(exc))
        except Exception:  # noqa: BLE001 — a handler bug must not kill the listener
            logger.exception("transport_handler_failed")
        finally:
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

        async def _go():
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

        return await asyncio.wait_for(_go(), timeout=timeout)


__all__ = ["Transport", "PeerInfo", "format_multiaddr", "parse_multiaddr", "read_frame", "safe_unpackb"]
