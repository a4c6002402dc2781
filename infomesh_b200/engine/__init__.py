"""GPU engines: sharded index, encoder / reranker / summariser wrappers, the hybrid search pipeline."""
