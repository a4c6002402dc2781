"""SimHash fingerprints + Hamming scan (csrc/dedup/simhash.cu, csrc/host/textproc.cpp) — K1.

``simhash_py`` is the pure-Python oracle with the reference's exact semantics
(reference infomesh/crawler/simhash.py:43-96); ``simhash_batch`` runs the CUDA kernel on a batch of
documents (bit-exact with the oracle), ``simhash_cpu`` the C++ host implementation.
"""
from __future__ import annotations

import ctypes
import hashlib
import re

import numpy as np

from infomesh_b200 import _native

_WORD_RE = re.compile(r"\w+", re.UNICODE)
SHINGLE_WIDTH = 3
HAMMING_THRESHOLD = 3


def words_of(text: str) -> list[str]:
    return _WORD_RE.findall(text.lower())


def simhash_py(text: str, width: int = SHINGLE_WIDTH) -> int:
    """Reference-semantics SimHash (pure Python oracle)."""
    words = words_of(text)
    if not words:
        return 0
    if len(words) < width:
        shingles = [" ".join(words)]
    else:
        shingles = [" ".join(words[i:i + width]) for i in range(len(words) - width + 1)]
    vec = [0] * 64
    for sh in shingles:
        h = int.from_bytes(hashlib.md5(sh.encode("utf-8")).digest()[:8], "big")  # noqa: S324
        for b in range(64):
            vec[b] += 1 if (h >> b) & 1 else -1
    fp = 0
    for b in range(64):
        if vec[b] >= 0:
            fp |= 1 << b
    return fp


def hamming(a: int, b: int) -> int:
    return bin((a ^ b) & 0xFFFFFFFFFFFFFFFF).count("1")


def normalize_batch(texts: list[str]):
    """Normalise documents for the kernels: lower-cased ``\\w+`` words joined by single spaces.

    Returns ``(text_bytes uint8[], word_start int64[], word_end int64[], doc_word_off int64[n+1])``.
    Uses the C++ ASCII fast path per document, Python's regex for non-ASCII input.
    """
    L = _native.lib() if _native.available() else None
    chunks: list[bytes] = []
    starts: list[np.ndarray] = []
    ends: list[np.ndarray] = []
    doc_off = [0]
    base = 0
    for t in texts:
        raw = t.encode("utf-8")
        done = False
        if L is not None and raw.isascii():
            out = ctypes.create_string_buffer(len(raw) + 1)
            cap = len(raw) // 2 + 2
            ws = np.empty(cap, dtype=np.int64)
            we = np.empty(cap, dtype=np.int64)
            olen = ctypes.c_longlong(0)
            n = L.im_normalize_words_ascii(raw, ctypes.c_longlong(len(raw)), out, ctypes.byref(olen),
                                           ws.ctypes.data_as(ctypes.c_void_p), we.ctypes.data_as(ctypes.c_void_p),
                                           ctypes.c_longlong(cap))
            if n >= 0:
                chunks.append(out.raw[:olen.value])
                starts.append(ws[:n] + base)
                ends.append(we[:n] + base)
                base += olen.value
                doc_off.append(doc_off[-1] + int(n))
                done = True
        if not done:
            words = [w.encode("utf-8") for w in words_of(t)]
            s = np.empty(len(words), dtype=np.int64)
            e = np.empty(len(words), dtype=np.int64)
            pos = base
            for i, w in enumerate(words):
                if i:
                    pos += 1
                s[i] = pos
                pos += len(w)
                e[i] = pos
            chunks.append(b" ".join(words))
            starts.append(s)
            ends.append(e)
            base = pos if words else base
            doc_off.append(doc_off[-1] + len(words))
    text = np.frombuffer(b"".join(chunks) + b"\0" * 8, dtype=np.uint8).copy()
    ws = np.concatenate(starts) if starts else np.zeros(0, np.int64)
    we = np.concatenate(ends) if ends else np.zeros(0, np.int64)
    return text, ws, we, np.asarray(doc_off, dtype=np.int64)


def simhash_cpu(texts: list[str], width: int = SHINGLE_WIDTH) -> np.ndarray:
    """C++ host SimHash (falls back to the Python oracle without the native library)."""
    if not _native.available():
        return np.asarray([simhash_py(t, width) for t in texts], dtype=np.uint64)
    L = _native.lib()
    L.im_simhash_cpu.restype = ctypes.c_ulonglong
    text, ws, we, off = normalize_batch(texts)
    out = np.zeros(len(texts), dtype=np.uint64)
    for i in range(len(texts)):
        a, b = int(off[i]), int(off[i + 1])
        out[i] = L.im_simhash_cpu(text.ctypes.data_as(ctypes.c_char_p), ws[a:b].ctypes.data_as(ctypes.c_void_p),
                                  we[a:b].ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(b - a),
                                  ctypes.c_int(width))
    return out


def simhash_batch(texts: list[str], width: int = SHINGLE_WIDTH, device="cuda"):
    """CUDA SimHash of a batch of documents -> int64 tensor (bit pattern of the uint64 fingerprints)."""
    import torch

    text, ws, we, off = normalize_batch(texts)
    return simhash_from_arrays(torch.from_numpy(text).to(device), torch.from_numpy(ws).to(device),
                               torch.from_numpy(we).to(device), torch.from_numpy(off).to(device), width)


def simhash_from_arrays(text, word_start, word_end, doc_word_off, width: int = SHINGLE_WIDTH):
    import torch

    n_docs = doc_word_off.numel() - 1
    out = torch.empty((n_docs,), device=text.device, dtype=torch.int64)
    L = _native.require()
    rc = L.im_simhash(_native.ptr(text), _native.ptr(word_start), _native.ptr(word_end), _native.ptr(doc_word_off),
                      ctypes.c_int(n_docs), ctypes.c_int(width), _native.ptr(out), _native.stream_ptr())
    _native.check(rc, "im_simhash")
    _native.count_launch()
    return out


def hamming_scan(table, probes, threshold: int = HAMMING_THRESHOLD, index_base: int = 0, best=None, n_table_dev=None):
    """For every probe: ``(distance, index)`` of the nearest table entry within ``threshold`` (else (-1,-1)).

    ``table`` / ``probes``: int64 CUDA tensors holding uint64 bit patterns.  ``best`` may be passed to
    accumulate over several table shards (packed ``dist << 32 | index`` with all-ones meaning "none").
    """
    import torch

    assert table.dtype == torch.int64 and probes.dtype == torch.int64
    if best is None:
        best = torch.full((probes.numel(),), -1, device=probes.device, dtype=torch.int64)  # 0xFFFF... pattern
    if table.numel() and probes.numel():
        L = _native.require()
        if n_table_dev is not None:     # fill level read on the device: the scan covers table[:min(len, *n_table_dev)]
            rc = L.im_hamming_scan_dev(_native.ptr(table), ctypes.c_longlong(table.numel()), ctypes.c_longlong(index_base),
                                       _native.ptr(probes), ctypes.c_int(probes.numel()), ctypes.c_int(threshold),
                                       _native.ptr(best), _native.ptr(n_table_dev), _native.stream_ptr())
        else:
            rc = L.im_hamming_scan(_native.ptr(table), ctypes.c_longlong(table.numel()), ctypes.c_longlong(index_base),
                                   _native.ptr(probes), ctypes.c_int(probes.numel()), ctypes.c_int(threshold),
                                   _native.ptr(best), _native.stream_ptr())
        _native.check(rc, "im_hamming_scan")
        _native.count_launch()
    return best


def dedup_resolve(all_fp, best_parts, threshold: int = HAMMING_THRESHOLD):
    """keep-mask uint8 ``[n]`` of a batch: no rank reported an indexed near-duplicate (``best_parts`` ``[P, n]`` packed scan
    results) and no earlier passage of the batch is within ``threshold`` bits.  One kernel, no host sync."""
    import torch

    n = all_fp.numel()
    bp = best_parts.reshape(-1, n).contiguous()
    keep = torch.empty((n,), device=all_fp.device, dtype=torch.uint8)
    L = _native.require()
    rc = L.im_dedup_resolve(_native.ptr(all_fp), _native.ptr(bp), ctypes.c_int(bp.shape[0]), ctypes.c_int(n), ctypes.c_int(threshold),
                            _native.ptr(keep), _native.stream_ptr())
    _native.check(rc, "im_dedup_resolve")
    _native.count_launch()
    return keep


def dedup_resolve_ref(all_fp, best_parts, threshold: int = HAMMING_THRESHOLD):
    """NumPy oracle of :func:`dedup_resolve`."""
    fp = all_fp.cpu().numpy().astype(np.uint64)
    bp = best_parts.reshape(-1, fp.size).cpu().numpy()
    keep = np.ones(fp.size, dtype=np.uint8)
    for j in range(fp.size):
        if (bp[:, j] != -1).any():
            keep[j] = 0
            continue
        x = fp[:j] ^ fp[j]
        if j and min(bin(int(v)).count("1") for v in x) <= threshold:
            keep[j] = 0
    return keep


def dedup_append(keep_slice, emb, fp, first_id: int, vectors, fingerprints, doc_ids, n_dev, counters):
    """Append the kept rows of this rank's slice to its shard at the device-side fill level ``n_dev`` (int64 ``[1]``);
    ``counters`` (int64 ``[3]``) accumulates seen / duplicates / overflow."""
    L = _native.require()
    bpr, H = emb.shape
    rc = L.im_dedup_append(_native.ptr(keep_slice), ctypes.c_int(bpr), _native.ptr(emb), ctypes.c_int(H), _native.ptr(fp),
                           ctypes.c_longlong(first_id), _native.ptr(vectors), _native.ptr(fingerprints), _native.ptr(doc_ids),
                           _native.ptr(n_dev), ctypes.c_longlong(vectors.shape[0]), _native.ptr(counters), _native.stream_ptr())
    _native.check(rc, "im_dedup_append")
    _native.count_launch()


def unpack_best(best):
    """packed -> (dist int64[-1 if none], index int64[-1 if none])"""
    none = best == -1
    dist = (best >> 32) & 0xFFFFFFFF
    idx = best & 0xFFFFFFFF
    return dist.masked_fill(none, -1), idx.masked_fill(none, -1)
