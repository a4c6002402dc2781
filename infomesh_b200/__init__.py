"""infomesh_b200 — a Blackwell-native (sm_100a) search / RAG node with InfoMesh's capabilities.

CPU plane (config, crawler, index, search, p2p, credits, trust, MCP, HTTP, CLI, SDK) is plain
Python; the hot paths (encoder / reranker / summariser GEMMs and attention, sharded vector search,
BM25 scoring, SimHash dedup, passage extraction) are hand-written CUDA kernels in ``csrc/``.
"""
__version__ = "0.1.0"
DISTRIBUTION = "infomesh-b200"      # the name this package is published and upgraded under (not the reference's `infomesh`)


import os as _os

if _os.environ.get("INFOMESH_B200_ALIAS", "") not in ("", "0"):      # run scripts written for the reference unchanged (compat.py)
    from infomesh_b200.compat import alias_as_infomesh as _alias

    _alias()
