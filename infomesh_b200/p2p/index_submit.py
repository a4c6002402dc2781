"""Split deployment: DMZ crawler nodes submit crawled pages to private indexer nodes
(``network.index_submit_peers``); the receiver checks the peer ACL and the Ed25519 signature over
``url:text_hash:raw_html_hash`` before indexing (reference infomesh/p2p/index_submit.py:40-292).
Transport: the signed P2P channel first; an HTTP POST of the same frame to ``<peer>/index/submit`` for peers given as
http(s) URLs."""
from __future__ import annotations

import time
from dataclasses import asdict
from typing import Any

from infomesh_b200.p2p.protocol import IndexSubmit, IndexSubmitAck, MessageType, encode_message
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


def _sign_blob(url: str, text_hash: str, raw_html_hash: str) -> bytes:
    return f"{url}:{text_hash}:{raw_html_hash}".encode()


class IndexSubmitSender:
    def __init__(self, config, key_pair=None, transport=None):
        self._config, self._key_pair, self._transport = config, key_pair, transport
        self._submit_peers = list(config.network.index_submit_peers)
        self._sent = self._errors = 0

    @property
    def submit_peers(self) -> list[str]:
        return self._submit_peers

    @property
    def stats(self) -> dict[str, int]:
        return {"sent": self._sent, "errors": self._errors}

    def build_submit(self, page, discovered_links: list[str] | None = None) -> IndexSubmit:
        pid, sig = "", b""
        if self._key_pair is not None:
            pid = self._key_pair.peer_id
            sig = self._key_pair.sign(_sign_blob(page.url, page.text_hash, page.raw_html_hash))
        return IndexSubmit(url=page.url, title=page.title, text=page.text, raw_html_hash=page.raw_html_hash,
                           text_hash=page.text_hash, language=page.language or "", crawled_at=time.time(), peer_id=pid,
                           signature=sig, discovered_links=list(discovered_links or []))

    def build_submit_message(self, page, discovered_links: list[str] | None = None) -> bytes:
        return encode_message(MessageType.INDEX_SUBMIT, asdict(self.build_submit(page, discovered_links)))

    def record_sent(self) -> None:
        self._sent += 1

    def record_error(self) -> None:
        self._errors += 1

    async def send_to_peers(self, message: bytes) -> int:
        import asyncio

        from infomesh_b200.p2p.protocol import decode_message

        ok = 0
        for peer in self._submit_peers:
            try:
                if peer.startswith(("http://", "https://")):
                    good = await asyncio.get_running_loop().run_in_executor(None, _http_post, peer, message)
                else:
                    if self._transport is None:
                        raise RuntimeError("no P2P transport for index-submit")
                    kind, body = decode_message(message)
                    reply = await self._transport.request(peer, kind, body, timeout=30.0)
                    good = bool(reply and reply[0] == MessageType.INDEX_SUBMIT_ACK and reply[1].get("success"))
            except Exception as exc:  # noqa: BLE001
                logger.warning("index_submit_send_failed", peer=peer, error=str(exc))
                good = False
            if good:
                self.record_sent()
                ok += 1
            else:
                self.record_error()
        return ok


def _http_post(peer: str, message: bytes) -> bool:
    import urllib.request

    req = urllib.request.Request(f"{peer.rstrip('/')}/index/submit", data=message,
                                 headers={"Content-Type": "application/x-msgpack"}, method="POST")
    with urllib.request.urlopen(req, timeout=30.0) as resp:  # noqa: S310 — operator-configured peer URL
        return resp.status < 400


class IndexSubmitReceiver:
    def __init__(self, config, store, vector_store=None, key_pair=None, key_registry=None):
        self._config, self._store, self._vector_store, self._key_pair = config, store, vector_store, key_pair
        self._acl = frozenset(config.network.peer_acl)
        self._keys = key_registry
        self._received = self._rejected = self._indexed = 0

    def bind_key_registry(self, key_registry) -> None:
        """Called by the P2P node once its transport exists: submissions are then checked against the verified keys."""
        self._keys = key_registry

    @property
    def stats(self) -> dict[str, int]:
        return {"received": self._received, "rejected": self._rejected, "indexed": self._indexed}

    def is_peer_allowed(self, peer_id: str) -> bool:
        return not self._acl or peer_id in self._acl

    def _ack(self, url: str, *, ok: bool, doc_id: int = 0, error: str = "") -> IndexSubmitAck:
        return IndexSubmitAck(url=url, doc_id=doc_id, success=ok, error=error,
                              peer_id=self._key_pair.peer_id if self._key_pair else "")

    def _reject(self, url: str, pid: str, reason: str) -> IndexSubmitAck:
        self._rejected += 1
        logger.warning("index_submit_rejected", peer_id=pid, url=url, reason=reason)
        return self._ack(url, ok=False, error=reason)

    def handle_submit(self, payload: dict[str, Any], sender: str | None = None) -> IndexSubmitAck:
        """``sender``: peer id the transport VERIFIED for the enclosing envelope ("" = unsigned frame, ``None`` = the
        caller has no transport identity, e.g. the localhost HTTP bridge).

        Checks, in order: the text really hashes to ``text_hash`` (a captured signature cannot be replayed with other
        text); with a ``peer_acl`` in force the claimed ``peer_id`` must be in it, must equal the transport-verified
        sender, and the Ed25519 signature over ``url:text_hash:raw_html_hash`` must verify under that peer's known key.
        In open mode a signature is still checked whenever the key is known."""
        from infomesh_b200.crawler.parser import ParsedPage
        from infomesh_b200.hashing import content_hash
        from infomesh_b200.p2p.keys import verify_with_public_key
        from infomesh_b200.services import index_document

        self._received += 1
        pid, url = str(payload.get("peer_id", "")), str(payload.get("url", ""))
        text_hash, raw_hash = str(payload.get("text_hash", "")), str(payload.get("raw_html_hash", ""))
        if content_hash(str(payload.get("text", ""))) != text_hash:
            return self._reject(url, pid, "text_hash_mismatch")
        if not self.is_peer_allowed(pid):
            return self._reject(url, pid, "peer_not_allowed")
        if sender is not None and pid and sender and sender != pid:
            return self._reject(url, pid, "peer_id_mismatch")
        pub = self._keys.get(pid) if self._keys is not None and pid else None
        sig = payload.get("signature") or b""
        if self._acl:
            # an ACL only means something if the claimed identity is proven end to end
            if sender is not None and sender != pid:
                return self._reject(url, pid, "unsigned_sender")
            if pub is None:
                return self._reject(url, pid, "unknown_key")
        if pub is not None and not verify_with_public_key(pub, _sign_blob(url, text_hash, raw_hash), bytes(sig)):
            return self._reject(url, pid, "bad_signature")
        page = ParsedPage(url=url, title=str(payload.get("title", "")), text=str(payload.get("text", "")),
                          raw_html_hash=raw_hash, text_hash=text_hash, language=str(payload.get("language", "")))
        try:
            doc_id = index_document(page, self._store, self._vector_store)
        except Exception as exc:  # noqa: BLE001
            logger.warning("index_submit_error", peer_id=pid, url=url, error=str(exc))
            return self._ack(url, ok=False, error=str(exc))
        self._indexed += 1
        return self._ack(url, ok=True, doc_id=doc_id or 0)

    def build_ack_message(self, ack: IndexSubmitAck) -> bytes:
        return encode_message(MessageType.INDEX_SUBMIT_ACK, asdict(ack))
