"""pytest configuration: the ``gpu`` marker and shared fixtures.

CPU tests run everywhere (``-m "not gpu"``); GPU tests need a B200 and the native library
(``python -m infomesh_b200.build``).  Numerics tests compare each CUDA kernel with a plain PyTorch fp32 reference.
"""
from __future__ import annotations

import asyncio
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("INFOMESH_HOME", "/tmp/infomesh_b200_test_home")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def tmp_data_dir(tmp_path):
    d = tmp_path / "infomesh"
    d.mkdir()
    return d


@pytest.fixture
def run_async():
    """Run a coroutine to completion (pytest-asyncio is not installed here)."""

    def runner(coro):
        return asyncio.run(coro)

    return runner


@pytest.fixture
def store():
    from infomesh_b200.index.local_store import LocalStore

    s = LocalStore()
    yield s
    s.close()


DOCS = [
    ("https://docs.python.org/3/library/asyncio.html", "asyncio — Asynchronous I/O",
     "asyncio is a library to write concurrent code using the async/await syntax. asyncio is used as a foundation "
     "for multiple Python asynchronous frameworks that provide high-performance network and web-servers."),
    ("https://www.python-httpx.org/", "HTTPX",
     "HTTPX is a fully featured HTTP client for Python 3, which provides sync and async APIs, and support for "
     "both HTTP/1.1 and HTTP/2. It has a requests-compatible API."),
    ("https://sqlite.org/fts5.html", "SQLite FTS5 Extension",
     "FTS5 is an SQLite virtual table module that provides full-text search functionality to database "
     "applications. The bm25 ranking function returns a value indicating how well a row matches the query."),
    ("https://example.com/rust/book", "The Rust Programming Language",
     "Rust is a systems programming language focused on safety, speed and concurrency. Ownership is its most "
     "unique feature and enables memory safety guarantees without a garbage collector."),
]


@pytest.fixture
def filled_store(store):
    from infomesh_b200.hashing import content_hash

    for url, title, text in DOCS:
        store.add_document(url, title, text, content_hash("<html>" + text), content_hash(text), language="en")
    return store
