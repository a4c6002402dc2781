"""Search layer: query orchestration, ranking helpers, passage extraction, RAG, NLP utilities."""
