"""Index three documents and search them — the minimal SDK round trip (no GPU, no network)."""
import tempfile

from infomesh_b200.sdk import InfoMeshClient

DOCS = [
    ("https://example.org/tmem", "Tensor memory on Blackwell", "Blackwell tensor cores accumulate into tensor memory (TMEM), 256 KB per SM. " * 4),
    ("https://example.org/kademlia", "Kademlia in one page", "Kademlia keeps k-buckets of contacts and performs iterative lookups with alpha parallelism. " * 4),
    ("https://example.org/bm25", "BM25 ranking", "BM25 scores a document by term frequency saturation and inverse document frequency. " * 4),
]

with tempfile.TemporaryDirectory() as d, InfoMeshClient(d) as client:
    for url, title, text in DOCS:
        client.add_document(url, title, text)
    for hit in client.search("tensor memory", limit=3):
        print(f"{hit.score:.3f}  {hit.title}  <{hit.url}>")
    print("suggest('Kad') ->", client.suggest("Kad"))
    print("stats ->", client.get_stats())
