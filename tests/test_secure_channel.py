"""p2p/secure_channel.py + its integration in p2p/transport.py: encrypted, mutually authenticated sessions."""
from __future__ import annotations

import asyncio

import pytest

from infomesh_b200.p2p import message_auth as MA
from infomesh_b200.p2p import secure_channel as SC
from infomesh_b200.p2p.keys import KeyPair
from infomesh_b200.p2p.protocol import MessageType, encode_message
from infomesh_b200.p2p.transport import Transport


async def _pair(server_kp, client_kp, expect=None):
    done: asyncio.Future = asyncio.get_running_loop().create_future()

    async def on_conn(r, w):
        try:
            assert await r.readexactly(4) == SC.MAGIC
            done.set_result((await SC.server_handshake(r, w, server_kp), r, w))
        except Exception as exc:  # noqa: BLE001
            if not done.done():
                done.set_exception(exc)

    srv = await asyncio.start_server(on_conn, "127.0.0.1", 0)
    port = srv.sockets[0].getsockname()[1]
    r, w = await asyncio.open_connection("127.0.0.1", port)
    cs = await SC.client_handshake(r, w, client_kp, expect_peer=expect)
    ss, sr, sw = await done
    return srv, (cs, r, w), (ss, sr, sw)


def test_handshake_authenticates_both_sides_and_encrypts_records():
    a, b = KeyPair.generate(), KeyPair.generate()

    async def go():
        srv, (cs, cr, cw), (ss, sr, sw) = await _pair(a, b, expect=a.peer_id)
        assert cs.remote_peer_id == a.peer_id and ss.remote_peer_id == b.peer_id
        secret = b"top secret query " * 10
        await cs.send(cw, secret)
        assert await ss.recv(sr) == secret
        await ss.send(sw, b"reply")
        assert await cs.recv(cr) == b"reply"
        rec = cs.seal(secret)
        assert secret not in rec and len(rec) == 4 + len(secret) + 16                 # ciphertext + tag, nothing in clear
        cw.close()
        sw.close()
        srv.close()

    asyncio.run(go())


def test_wrong_identity_tampering_and_replay_are_rejected():
    a, b, mallory = KeyPair.generate(), KeyPair.generate(), KeyPair.generate()

    async def go():
        with pytest.raises(SC.HandshakeError, match="expected"):
            await _pair(mallory, b, expect=a.peer_id)                                    # dialled a, reached mallory
        srv, (cs, cr, cw), (ss, sr, sw) = await _pair(a, b)
        rec = cs.seal(b"hello")
        with pytest.raises(SC.HandshakeError):
            ss.open(bytes([rec[4] ^ 1]) + rec[5:])                                        # flipped ciphertext bit
        assert ss.open(rec[4:]) == b"hello"
        with pytest.raises(SC.HandshakeError):
            ss.open(rec[4:])                                                              # replayed record: counter moved on
        cw.close()
        sw.close()
        srv.close()

    asyncio.run(go())


def test_transport_uses_the_encrypted_channel_and_authenticates_senders():
    a, b = KeyPair.generate(), KeyPair.generate()

    async def go():
        ta, tb = Transport(a), Transport(b)
        seen = {}

        async def on_ping(payload, peer):
            seen["peer"] = peer.peer_id
            return MessageType.PONG, {"echo": payload.get("n"), "peer_id": a.peer_id}

        async def on_submit(payload, peer):                        # a WRITE type: must arrive authenticated
            seen["submit_from"] = peer.peer_id
            return MessageType.INDEX_SUBMIT_ACK, {"ok": True}

        ta.register(MessageType.PING, on_ping)
        ta.register(MessageType.INDEX_SUBMIT, on_submit)
        await ta.listen("127.0.0.1", 0)
        kind, body = await tb.request(("127.0.0.1", ta.port), MessageType.PING, {"n": 7}, expect_peer=a.peer_id)
        assert kind == MessageType.PONG and body["echo"] == 7 and seen["peer"] == b.peer_id
        assert tb.encrypted_out == 1 and ta.encrypted_in == 1 and b.peer_id in ta.keys and a.peer_id in tb.keys
        await tb.request(("127.0.0.1", ta.port), MessageType.INDEX_SUBMIT, {"url": "u"})
        assert seen["submit_from"] == b.peer_id
        with pytest.raises(MA.VerificationError):
            await tb.request(("127.0.0.1", ta.port), MessageType.PING, {}, expect_peer=KeyPair.generate().peer_id)
        # a plaintext-only peer is still served (and signed envelopes still authenticate it) ...
        legacy = Transport(KeyPair.generate(), encrypt=False)
        kind, _ = await legacy.request(("127.0.0.1", ta.port), MessageType.PING, {"n": 1})
        assert kind == MessageType.PONG and legacy.encrypted_out == 0
        # ... unless the node insists on encryption
        ta.require_encrypted = True
        kind, body = await legacy.request(("127.0.0.1", ta.port), MessageType.PING, {"n": 1})
        assert kind == MessageType.ERROR and "encrypted" in body["error"]
        await ta.close()

    asyncio.run(go())


def test_encrypted_client_falls_back_for_a_peer_without_the_handshake():
    a, b = KeyPair.generate(), KeyPair.generate()

    async def go():
        old = Transport(a, encrypt=False)                          # e.g. a round-1 node

        async def on_ping(payload, peer):
            return MessageType.PONG, {"ok": True}

        old.register(MessageType.PING, on_ping)
        await old.listen("127.0.0.1", 0)
        new = Transport(b)
        kind, body = await new.request(("127.0.0.1", old.port), MessageType.PING, {})
        assert kind == MessageType.PONG and ("127.0.0.1", old.port) in new._plain_peers and new.encrypted_out == 0
        strict = Transport(KeyPair.generate(), require_encrypted=True)
        with pytest.raises((ConnectionError, asyncio.IncompleteReadError, SC.HandshakeError)):
            await strict.request(("127.0.0.1", old.port), MessageType.PING, {})
        await old.close()

    asyncio.run(go())


def test_unsigned_write_frames_are_refused_on_plaintext_connections():
    async def go():
        t = Transport(KeyPair.generate())

        async def on_submit(payload, peer):
            return MessageType.INDEX_SUBMIT_ACK, {"ok": True}

        t.register(MessageType.INDEX_SUBMIT, on_submit)
        await t.listen("127.0.0.1", 0)
        r, w = await asyncio.open_connection("127.0.0.1", t.port)
        w.write(encode_message(MessageType.INDEX_SUBMIT, {"url": "u"}))
        await w.drain()
        from infomesh_b200.p2p.protocol import decode_message
        from infomesh_b200.p2p.transport import read_frame

        kind, body = decode_message(await read_frame(r))
        assert kind == MessageType.ERROR and "unsigned" in body["error"]
        w.close()
        await t.close()

    asyncio.run(go())


def test_encrypted_sessions_are_reused_and_redialled_after_a_drop():
    """One handshake serves many requests to the same peer; a session the listener dropped is re-established transparently;
    fire-and-forget messages never share a pooled session."""
    import asyncio

    from infomesh_b200.p2p import transport as T
    from infomesh_b200.p2p.keys import KeyPair
    from infomesh_b200.p2p.protocol import MessageType

    async def go():
        a, b = T.Transport(KeyPair.generate()), T.Transport(KeyPair.generate())
        seen = []

        async def ping(payload, peer):
            seen.append((payload.get("n"), peer.peer_id))
            return MessageType.PONG, {"n": payload.get("n")}

        b.register(MessageType.PING, ping)
        port = await b.listen("127.0.0.1", 0)
        addr = ("127.0.0.1", port)
        try:
            for n in range(5):
                kind, body = await a.request(addr, MessageType.PING, {"n": n})
                assert kind == MessageType.PONG and body["n"] == n
            assert a.handshakes_out == 1 and a.encrypted_out == 5 and b.encrypted_in == 5
            assert {pid for _, pid in seen} == {a.peer_id}
            # the listener drops every inbound connection (restart / idle timeout): the next request dials again, once
            for w in list(b._inbound):
                w.close()
            await asyncio.sleep(0.05)
            kind, body = await a.request(addr, MessageType.PING, {"n": 99})
            assert body["n"] == 99 and a.handshakes_out == 2
            # concurrent requests to one peer are serialised on its session, not interleaved
            outs = await asyncio.gather(*(a.request(addr, MessageType.PING, {"n": 100 + i}) for i in range(8)))
            assert sorted(o[1]["n"] for o in outs) == list(range(100, 108)) and a.handshakes_out == 2
            # fire-and-forget uses a private connection
            assert await a.request(addr, MessageType.PING, {"n": -1}, expect_reply=False) is None
            assert a.handshakes_out == 3
            kind, body = await a.request(addr, MessageType.PING, {"n": 7})
            assert body["n"] == 7 and a.handshakes_out == 3
            # identity pinning also applies to a pooled session
            import pytest as _pt

            from infomesh_b200.p2p import message_auth as MA
            with _pt.raises((MA.VerificationError, Exception)):
                await a.request(addr, MessageType.PING, {"n": 1}, expect_peer="f" * 40)
        finally:
            await a.close()
            await asyncio.wait_for(b.close(), timeout=5)          # open inbound sessions must not hold the listener's shutdown

    asyncio.run(go())
