"""Starter index for cold-start nodes: locate the ``starter.infomesh-snapshot`` release asset (metadata cached for an
hour), stream it to the data dir with progress reporting, and tell callers when a node is empty enough (< 10
documents) to want it (reference infomesh/index/starter.py:33-250)."""
from __future__ import annotations

import asyncio
import contextlib
import json
import time
from pathlib import Path
from typing import Any, Callable

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

GITHUB_REPO = "dotnetpower/infomesh"
RELEASES_API = f"https://api.github.com/repos/{GITHUB_REPO}/releases"
SNAPSHOT_ASSET_NAME = "starter.infomesh-snapshot"
_CACHE_FILE = "starter_meta_cache.json"
_CACHE_TTL = 3600
_REQUEST_TIMEOUT = 10.0
_DOWNLOAD_TIMEOUT = 600.0
_CHUNK_SIZE = 65536


class StarterAssetInfo:
    __slots__ = ("download_url", "size_bytes", "release_tag", "created_at")

    def __init__(self, download_url: str, size_bytes: int, release_tag: str, created_at: str):
        self.download_url, self.size_bytes, self.release_tag, self.created_at = download_url, size_bytes, release_tag, created_at

    @property
    def size_mb(self) -> float:
        return self.size_bytes / 2 ** 20


def pick_asset(releases: list[dict[str, Any]]) -> StarterAssetInfo | None:
    """First release (newest first) that carries the snapshot asset."""
    for rel in releases if isinstance(releases, list) else []:
        for asset in rel.get("assets", []) or []:
            if asset.get("name") == SNAPSHOT_ASSET_NAME and asset.get("browser_download_url"):
                return StarterAssetInfo(asset["browser_download_url"], int(asset.get("size", 0)), str(rel.get("tag_name", "")),
                                        str(asset.get("created_at", "")))
    return None


def _read_cache(cache_dir: Path) -> StarterAssetInfo | None:
    try:
        d = json.loads((Path(cache_dir) / _CACHE_FILE).read_text("utf-8"))
        if time.time() - d.get("ts", 0) > _CACHE_TTL:
            return None
        return StarterAssetInfo(d["url"], int(d["size"]), d["tag"], d.get("created_at", ""))
    except Exception:  # noqa: BLE001
        return None


def _write_cache(cache_dir: Path, info: StarterAssetInfo) -> None:
    with contextlib.suppress(Exception):
        (Path(cache_dir) / _CACHE_FILE).write_text(json.dumps({"ts": time.time(), "url": info.download_url, "size": info.size_bytes,
                                                                "tag": info.release_tag, "created_at": info.created_at}), encoding="utf-8")


async def find_starter_asset(*, cache_dir: Path | None = None) -> StarterAssetInfo | None:
    if cache_dir is not None:
        hit = _read_cache(cache_dir)
        if hit is not None:
            return hit
    try:
        import httpx

        async with httpx.AsyncClient(timeout=_REQUEST_TIMEOUT) as client:
            resp = await client.get(RELEASES_API, headers={"Accept": "application/vnd.github+json"})
            resp.raise_for_status()
            info = pick_asset(resp.json())
    except Exception as exc:  # noqa: BLE001
        logger.debug("starter_lookup_failed", error=str(exc))
        return None
    if info is not None and cache_dir is not None:
        _write_cache(cache_dir, info)
    return info


async def download_starter_snapshot(data_dir: Path, *, progress_callback: Callable[[int, int], None] | None = None) -> Path | None:
    data_dir = Path(data_dir)
    asset = await find_starter_asset(cache_dir=data_dir)
    if asset is None:
        return None
    dest = data_dir / SNAPSHOT_ASSET_NAME
    if dest.exists() and dest.stat().st_size == asset.size_bytes:
        return dest
    tmp = dest.with_suffix(".tmp")
    try:
        import httpx

        async with httpx.AsyncClient(follow_redirects=True, timeout=httpx.Timeout(_DOWNLOAD_TIMEOUT, connect=10.0)) as client, \
                client.stream("GET", asset.download_url) as resp:
            resp.raise_for_status()
            got = 0
            with open(tmp, "wb") as f:
                async for chunk in resp.aiter_bytes(chunk_size=_CHUNK_SIZE):
                    f.write(chunk)
                    got += len(chunk)
                    if progress_callback is not None:
                        progress_callback(got, asset.size_bytes)
        tmp.replace(dest)
        return dest
    except Exception as exc:  # noqa: BLE001
        logger.warning("starter_download_failed", error=str(exc))
        tmp.unlink(missing_ok=True)
        return None


def needs_starter(index_doc_count: int) -> bool:
    return index_doc_count < 10


def download_starter_sync(data_dir: Path, *, progress_callback: Callable[[int, int], None] | None = None) -> Path | None:
    return asyncio.run(download_starter_snapshot(data_dir, progress_callback=progress_callback))
