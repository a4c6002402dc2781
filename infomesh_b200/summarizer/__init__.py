"""Summarisation: pluggable LLM backends (HTTP runtimes + the in-process sm_100a T5), verification, peer serving."""
