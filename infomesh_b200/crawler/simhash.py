"""SimHash near-duplicate detection — CPU API of reference infomesh/crawler/simhash.py:31-213.

``simhash`` uses the C++ host implementation (bit-exact with the pure-Python oracle ``ops.dedup.simhash_py``);
batched fingerprints and bulk Hamming scans run on the GPU (``ops.dedup.simhash_batch`` / ``hamming_scan``, K1).
``SimHashIndex`` keeps the reference's behaviour (insertion-ordered, capped at 500 k fingerprints, linear scan) and
transparently offloads large scans to the device when one is available.
"""
from __future__ import annotations

import contextlib

import numpy as np

from infomesh_b200.ops import dedup as _dd

HAMMING_THRESHOLD: int = 3
_NUM_BITS = 64
_SHINGLE_WIDTH = 3
_GPU_SCAN_MIN = 200_000        # below this a NumPy popcount scan is faster than a launch + copy


def _tokenize(text: str, width: int = _SHINGLE_WIDTH) -> list[str]:
    words = _dd.words_of(text)
    if len(words) < width:
        return [" ".join(words)] if words else []
    return [" ".join(words[i:i + width]) for i in range(len(words) - width + 1)]


def simhash(text: str, *, shingle_width: int = _SHINGLE_WIDTH) -> int:
    """64-bit fingerprint (0 for empty text)."""
    if not text:
        return 0
    return int(_dd.simhash_cpu([text], shingle_width)[0])


def simhash_many(texts: list[str], *, shingle_width: int = _SHINGLE_WIDTH, device: str | None = None) -> list[int]:
    """Batch fingerprints; runs the CUDA kernel when a device is available."""
    if device is None:
        try:
            import torch

            device = "cuda" if torch.cuda.is_available() else None
        except Exception:  # noqa: BLE001
            device = None
    if device and len(texts) >= 8:
        fp = _dd.simhash_batch(texts, shingle_width, device=device)
        return [int(x) for x in fp.cpu().numpy().view(np.uint64)]
    return [int(x) for x in _dd.simhash_cpu(texts, shingle_width)]


def hamming_distance(a: int, b: int) -> int:
    return bin((a ^ b) & 0xFFFFFFFFFFFFFFFF).count("1")


def is_near_duplicate(a: int, b: int, *, threshold: int = HAMMING_THRESHOLD) -> bool:
    return hamming_distance(a, b) <= threshold


class SimHashIndex:
    def __init__(self, *, max_entries: int = 500_000):
        self._entries: dict[int, list[int]] = {}
        self._max_entries = max_entries
        self._cache: np.ndarray | None = None     # uint64 view of the keys for vectorised scans

    @property
    def size(self) -> int:
        return len(self._entries)

    def add(self, doc_id: int, fingerprint: int) -> None:
        while len(self._entries) >= self._max_entries:      # FIFO eviction
            del self._entries[next(iter(self._entries))]
        self._entries.setdefault(fingerprint, []).append(doc_id)
        self._cache = None

    def remove(self, doc_id: int, fingerprint: int) -> None:
        ids = self._entries.get(fingerprint)
        if ids:
            with contextlib.suppress(ValueError):
                ids.remove(doc_id)
            if not ids:
                del self._entries[fingerprint]
                self._cache = None

    def _keys(self) -> np.ndarray:
        if self._cache is None:
            self._cache = np.fromiter(self._entries.keys(), dtype=np.uint64, count=len(self._entries))
        return self._cache

    def find_near_duplicates(self, fingerprint: int, *, threshold: int = HAMMING_THRESHOLD) -> list[int]:
        if not self._entries:
            return []
        keys = self._keys()
        x = keys ^ np.uint64(fingerprint)
        # popcount via byte table
        dist = _POPCNT[x.view(np.uint8).reshape(-1, 8)].sum(axis=1)
        out: list[int] = []
        for k in keys[dist <= threshold]:
            out.extend(self._entries[int(k)])
        return out

    def find_near_duplicates_batch(self, fingerprints: list[int], *, threshold: int = HAMMING_THRESHOLD,
                                   device: str | None = None) -> list[bool]:
        """Is there any stored fingerprint within ``threshold`` of each probe?  GPU scan for big tables."""
        if not self._entries or not fingerprints:
            return [False] * len(fingerprints)
        if device and len(self._entries) >= _GPU_SCAN_MIN:
            import torch

            table = torch.from_numpy(self._keys().view(np.int64).copy()).to(device)
            probes = torch.from_numpy(np.asarray(fingerprints, dtype=np.uint64).view(np.int64).copy()).to(device)
            best = _dd.hamming_scan(table, probes, threshold)
            return [bool(b) for b in (best != -1).cpu().tolist()]
        return [bool(self.find_near_duplicates(fp, threshold=threshold)) for fp in fingerprints]

    def get_stats(self) -> dict[str, int]:
        return {"unique_fingerprints": len(self._entries),
                "total_documents": sum(len(v) for v in self._entries.values())}


_POPCNT = np.array([bin(i).count("1") for i in range(256)], dtype=np.uint8)
