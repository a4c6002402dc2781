"""Compression codecs (zstd through the system libzstd, zlib fallback)."""
