"""Content change detection between crawls and WARC/1.0 export (reference infomesh/crawler/diff.py:15-152)."""
from __future__ import annotations

import hashlib
import time
from dataclasses import dataclass, field
from datetime import UTC, datetime


@dataclass
class ContentDiff:
    url: str
    has_changed: bool
    added_lines: list[str] = field(default_factory=list)
    removed_lines: list[str] = field(default_factory=list)
    change_ratio: float = 0.0
    old_length: int = 0
    new_length: int = 0


def compute_diff(old_text: str, new_text: str, url: str = "") -> ContentDiff:
    """Set-based line diff; ratio = changed lines / union of lines."""
    if old_text == new_text:
        return ContentDiff(url, False, old_length=len(old_text), new_length=len(new_text))
    old, new = set(old_text.splitlines()), set(new_text.splitlines())
    added = [ln for ln in new - old if ln.strip()]
    removed = [ln for ln in old - new if ln.strip()]
    ratio = (len(added) + len(removed)) / max(len(old | new), 1)
    return ContentDiff(url, True, added[:100], removed[:100], round(ratio, 3), len(old_text), len(new_text))


def _simple_uuid(url: str, ts: float) -> str:
    h = hashlib.sha256(f"{url}:{ts}".encode()).hexdigest()
    return f"{h[:8]}-{h[8:12]}-{h[12:16]}-{h[16:20]}-{h[20:32]}"


def export_warc_record(url: str, text: str, crawled_at: float) -> str:
    date = datetime.fromtimestamp(crawled_at, tz=UTC).strftime("%Y-%m-%dT%H:%M:%SZ")
    head = ("WARC/1.0\r\nWARC-Type: conversion\r\n"
            f"WARC-Target-URI: {url}\r\nWARC-Date: {date}\r\n"
            f"WARC-Record-ID: <urn:uuid:{_simple_uuid(url, crawled_at)}>\r\n"
            f"Content-Length: {len(text.encode('utf-8'))}\r\nContent-Type: text/plain\r\n\r\n")
    return head + text + "\r\n\r\n"


def export_warc_file(documents: list[dict[str, object]]) -> str:
    parts = ["WARC/1.0\r\nWARC-Type: warcinfo\r\n"
             f"WARC-Date: {time.strftime('%Y-%m-%dT%H:%M:%SZ', time.gmtime())}\r\n"
             "Content-Type: application/warc-fields\r\nContent-Length: 0\r\n\r\n\r\n"]
    for d in documents:
        url, text, ts = str(d.get("url", "")), str(d.get("text", "")), d.get("crawled_at", 0)
        if url and text:
            parts.append(export_warc_record(url, text, float(ts) if isinstance(ts, (int, float, str)) else 0.0))
    return "".join(parts)
