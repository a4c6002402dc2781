"""Portable index snapshots — file format of reference infomesh/index/snapshot.py:10-263:

    [u32 BE header_len][zstd(JSON metadata)][zstd(msgpack list of documents)]

Guards: 1 GiB file, 10 MiB header, 100 k documents.  Import dedups by ``text_hash`` (through the store's UNIQUE
constraints) and can feed a vector store.  (Readable by / from the reference when libzstd is present.)
"""
from __future__ import annotations

import json
import struct
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import msgpack

from infomesh_b200.compression.zstd import LEVEL_SNAPSHOT, Compressor
from infomesh_b200.index.local_store import LocalStore
from infomesh_b200.p2p.protocol import _SAFE_UNPACK
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

SNAPSHOT_EXTENSION = ".infomesh-snapshot"
_FORMAT_VERSION = 1
_MAX_SNAPSHOT_FILE_BYTES = 1024 * 1024 * 1024
_MAX_HEADER_LEN = 10 * 1024 * 1024
_MAX_SNAPSHOT_DOCUMENTS = 100_000
_DOC_FIELDS = ("url", "title", "text", "raw_html_hash", "text_hash")


@dataclass(frozen=True)
class SnapshotStats:
    total_documents: int
    exported: int      # export: written; import: newly added
    skipped: int       # import: duplicates
    file_size_bytes: int
    elapsed_ms: float


def export_snapshot(store: LocalStore, output_path: Path | str, *, compression_level: int = LEVEL_SNAPSHOT
                    ) -> SnapshotStats:
    t0 = time.monotonic()
    out = Path(output_path)
    comp = Compressor(level=compression_level)
    docs = store.export_documents()
    header = comp.compress(json.dumps({"format_version": _FORMAT_VERSION, "created_at": time.time(),
                                       "document_count": len(docs)}).encode())
    body = comp.compress(msgpack.packb(docs, use_bin_type=True))
    out.parent.mkdir(parents=True, exist_ok=True)
    tmp = out.with_name(out.name + ".part")
    with open(tmp, "wb") as fh:
        fh.write(struct.pack(">I", len(header)))
        fh.write(header)
        fh.write(body)
    tmp.replace(out)
    size = out.stat().st_size
    logger.info("snapshot_exported", documents=len(docs), file_size=size, path=str(out))
    return SnapshotStats(len(docs), len(docs), 0, size, (time.monotonic() - t0) * 1000)


def _read_header(fh) -> bytes:
    prefix = fh.read(4)
    if len(prefix) != 4:
        raise ValueError("Snapshot file is too small to contain a header")
    (n,) = struct.unpack(">I", prefix)
    if n > _MAX_HEADER_LEN:
        raise ValueError(f"Snapshot header too large: {n} bytes (max {_MAX_HEADER_LEN})")
    blob = fh.read(n)
    if len(blob) != n:
        raise ValueError("Snapshot header truncated")
    return blob


def read_snapshot_metadata(snapshot_path: Path | str) -> dict[str, Any]:
    with open(snapshot_path, "rb") as fh:
        blob = _read_header(fh)
    return dict(json.loads(Compressor(level=LEVEL_SNAPSHOT).decompress(blob)))


def import_snapshot(store: LocalStore, snapshot_path: Path | str, *, vector_store: Any | None = None
                    ) -> SnapshotStats:
    t0 = time.monotonic()
    path = Path(snapshot_path)
    size = path.stat().st_size
    if size > _MAX_SNAPSHOT_FILE_BYTES:
        raise ValueError(f"Snapshot file too large: {size} bytes (max {_MAX_SNAPSHOT_FILE_BYTES})")
    comp = Compressor(level=LEVEL_SNAPSHOT)
    with open(path, "rb") as fh:
        header_blob = _read_header(fh)
        body_blob = fh.read()
    meta = json.loads(comp.decompress(header_blob))
    version = meta.get("format_version", 0)
    if version > _FORMAT_VERSION:
        raise ValueError(f"Snapshot format version {version} is newer than supported ({_FORMAT_VERSION})")
    declared = meta.get("document_count", 0)
    if isinstance(declared, int) and declared > _MAX_SNAPSHOT_DOCUMENTS:
        raise ValueError(f"Snapshot document count too large: {declared} (max {_MAX_SNAPSHOT_DOCUMENTS})")
    docs = msgpack.unpackb(comp.decompress(body_blob, max_output_size=_MAX_SNAPSHOT_FILE_BYTES), raw=False,
                           **_SAFE_UNPACK)
    if not isinstance(docs, list):
        raise ValueError("Snapshot document payload must be a list")
    if len(docs) > _MAX_SNAPSHOT_DOCUMENTS:
        raise ValueError(f"Snapshot document count too large: {len(docs)} (max {_MAX_SNAPSHOT_DOCUMENTS})")
    added = skipped = 0
    for d in docs:
        if not isinstance(d, dict) or any(not isinstance(d.get(k), str) for k in _DOC_FIELDS):
            skipped += 1
            continue
        doc_id = store.add_document(d["url"], d["title"], d["text"], d["raw_html_hash"], d["text_hash"],
                                    language=d.get("language"))
        if doc_id is None:
            skipped += 1
            continue
        added += 1
        if vector_store is not None and hasattr(vector_store, "add_document"):
            vector_store.add_document(doc_id=doc_id, url=d["url"], title=d["title"], text=d["text"],
                                      language=d.get("language"))
    logger.info("snapshot_imported", imported=added, skipped=skipped, total=len(docs), path=str(path))
    return SnapshotStats(len(docs), added, skipped, size, (time.monotonic() - t0) * 1000)
