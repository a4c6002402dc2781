"""Shared lazy client handling for the framework adapters."""
from __future__ import annotations

from typing import Any


class ClientBacked:
    def __init__(self, data_dir: str = "~/.infomesh", *, client: Any | None = None, gpu: bool = False):
        self._data_dir, self._client, self._gpu = data_dir, client, gpu

    def _ensure_client(self):
        if self._client is None:
            from infomesh_b200.sdk.client import InfoMeshClient

            self._client = InfoMeshClient(data_dir=self._data_dir, gpu=self._gpu)
        return self._client

    def close(self) -> None:
        if self._client is not None:
            self._client.close()
            self._client = None


def result_meta(r: Any) -> dict[str, Any]:
    return {"title": r.title, "url": r.url, "score": r.score, "source": "infomesh"}
