"""Adapters for LangChain, LlamaIndex and Haystack (duck-typed: none of the frameworks is a dependency)."""
