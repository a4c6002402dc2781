"""infomesh_b200 — a Blackwell-native (sm_100a) search / RAG node with InfoMesh's capabilities.

CPU plane (config, crawler, index, search, p2p, credits, trust, MCP, HTTP, CLI, SDK) is plain
Python; the hot paths (encoder / reranker / summariser GEMMs and attention, sharded vector search,
BM25 scoring, SimHash dedup, passage extraction) are hand-written CUDA kernels in ``csrc/``.
"""
__version__ = "0.1.0"
DISTRIBUTION = "infomesh-b200"      # the name this package is published and upgraded under (not the reference's `infomesh`)
