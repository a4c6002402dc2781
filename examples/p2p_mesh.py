"""Three peers on loopback: B publishes a document's keywords to the DHT, C finds and queries B through the router."""
import asyncio
import tempfile
from dataclasses import replace
from pathlib import Path

from infomesh_b200.config import Config
from infomesh_b200.p2p.node import InfoMeshNode


def make(dirpath: Path, boot: list[str], **kw) -> InfoMeshNode:
    base = Config()
    cfg = replace(base, node=replace(base.node, data_dir=dirpath, listen_address="127.0.0.1", listen_port=0),
                  network=replace(base.network, bootstrap_nodes=boot or ["/ip4/127.0.0.1/tcp/1"], bootstrap_dns=False, bootstrap_github=False))
    return InfoMeshNode(cfg, pow_difficulty=8, enable_mdns=False, **kw)


async def answer(query: str, limit: int):
    return [{"url": "https://b.example/doc", "title": "B's document", "snippet": f"about {query}", "score": 0.9, "doc_id": 1}]


with tempfile.TemporaryDirectory() as d:
    a = make(Path(d) / "a", [])
    a.start()
    b, c = make(Path(d) / "b", a.listen_addrs, local_search_fn=answer), make(Path(d) / "c", a.listen_addrs)
    b.start()
    c.start()

    async def demo():
        await b.publish_document_to_network(1, "https://b.example/doc", "B's document", "blackwell tensor memory kernels")
        return await c.search_network("blackwell kernels", ["blackwell", "kernels"], 5)

    print("peers of a:", [p[:12] for p in a.get_connected_peers()])
    print("c's network search ->", asyncio.run(demo()))
    for n in (c, b, a):
        n.stop()
