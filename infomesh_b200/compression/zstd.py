"""zstd compression without the ``zstandard`` wheel: the system ``libzstd.so.1`` is driven through ctypes.

API parity with reference infomesh/compression/zstd.py:19-122 (``Compressor.compress/decompress/
compress_text/decompress_text``, levels 3/12/19, 100 MB decompression-bomb cap, optional dictionary,
``train_dictionary``).  If libzstd cannot be loaded, frames are written with zlib behind a 4-byte tag so data
written by either codec stays readable by this module (real zstd frames need libzstd to be read back).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import zlib

LEVEL_REALTIME = 3
LEVEL_SNAPSHOT = 12
LEVEL_ARCHIVE = 19
DEFAULT_MAX_OUTPUT_SIZE = 100 * 1024 * 1024
_ZSTD_MAGIC = b"\x28\xb5\x2f\xfd"
_ZLIB_TAG = b"IMZ1"
_CONTENTSIZE_UNKNOWN = 2**64 - 1
_CONTENTSIZE_ERROR = 2**64 - 2


class ZstdError(Exception):
    """Raised for corrupt input or output above the size limit."""


def _load():
    for name in ("libzstd.so.1", ctypes.util.find_library("zstd") or ""):
        if not name:
            continue
        try:
            lib = ctypes.CDLL(name)
        except OSError:
            continue
        lib.ZSTD_compressBound.restype = ctypes.c_size_t
        lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        lib.ZSTD_compress.restype = ctypes.c_size_t
        lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        lib.ZSTD_decompress.restype = ctypes.c_size_t
        lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        lib.ZSTD_isError.restype = ctypes.c_uint
        lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
        lib.ZSTD_getErrorName.restype = ctypes.c_char_p
        lib.ZSTD_getErrorName.argtypes = [ctypes.c_size_t]
        lib.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
        lib.ZSTD_getFrameContentSize.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        for fn in ("ZSTD_createCCtx", "ZSTD_createDCtx"):
            getattr(lib, fn).restype = ctypes.c_void_p
        lib.ZSTD_freeCCtx.argtypes = [ctypes.c_void_p]
        lib.ZSTD_freeDCtx.argtypes = [ctypes.c_void_p]
        lib.ZSTD_compress_usingDict.restype = ctypes.c_size_t
        lib.ZSTD_compress_usingDict.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                                ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        lib.ZSTD_decompress_usingDict.restype = ctypes.c_size_t
        lib.ZSTD_decompress_usingDict.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                                  ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        try:
            lib.ZDICT_trainFromBuffer.restype = ctypes.c_size_t
            lib.ZDICT_trainFromBuffer.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                                  ctypes.POINTER(ctypes.c_size_t), ctypes.c_uint]
        except AttributeError:
            pass
        return lib
    return None


_LIB = _load()


def have_zstd() -> bool:
    return _LIB is not None


def _check(code: int) -> int:
    if _LIB.ZSTD_isError(code):
        raise ZstdError(_LIB.ZSTD_getErrorName(code).decode())
    return code


class Compressor:
    """Reusable (de)compressor; ``dict_data`` enables dictionary mode for many small similar documents."""

    def __init__(self, level: int = LEVEL_REALTIME, *, dict_data: bytes | None = None):
        self._level = int(level)
        self._dict = bytes(dict_data) if dict_data else None

    @property
    def level(self) -> int:
        return self._level

    def compress(self, data: bytes) -> bytes:
        data = bytes(data)
        if _LIB is None:
            return _ZLIB_TAG + zlib.compress(data, min(9, max(1, self._level // 2)))
        bound = _LIB.ZSTD_compressBound(len(data))
        dst = ctypes.create_string_buffer(bound)
        if self._dict:
            cctx = _LIB.ZSTD_createCCtx()
            try:
                n = _check(_LIB.ZSTD_compress_usingDict(cctx, dst, bound, data, len(data), self._dict,
                                                         len(self._dict), self._level))
            finally:
                _LIB.ZSTD_freeCCtx(cctx)
        else:
            n = _check(_LIB.ZSTD_compress(dst, bound, data, len(data), self._level))
        return dst.raw[:n]

    def decompress(self, data: bytes, *, max_output_size: int = DEFAULT_MAX_OUTPUT_SIZE) -> bytes:
        data = bytes(data)
        if data[:4] == _ZLIB_TAG:
            d = zlib.decompressobj()
            out = d.decompress(data[4:], max_output_size + 1)
            if len(out) > max_output_size or d.unconsumed_tail:
                raise ZstdError(f"decompressed output exceeds limit of {max_output_size}")
            return out
        if data[:4] != _ZSTD_MAGIC:
            raise ZstdError("not a zstd frame")
        if _LIB is None:
            raise ZstdError("libzstd unavailable: cannot read a zstd frame")
        size = _LIB.ZSTD_getFrameContentSize(data, len(data))
        if size == _CONTENTSIZE_ERROR:
            raise ZstdError("corrupt zstd frame header")
        if size == _CONTENTSIZE_UNKNOWN:
            size = max_output_size  # streaming frame without a size field: cap at the limit
        if size > max_output_size:
            raise ZstdError(f"decompressed {size} bytes exceeds limit of {max_output_size}")
        dst = ctypes.create_string_buffer(max(int(size), 1))
        if self._dict:
            dctx = _LIB.ZSTD_createDCtx()
            try:
                n = _check(_LIB.ZSTD_decompress_usingDict(dctx, dst, int(size), data, len(data), self._dict,
                                                           len(self._dict)))
            finally:
                _LIB.ZSTD_freeDCtx(dctx)
        else:
            n = _check(_LIB.ZSTD_decompress(dst, int(size), data, len(data)))
        return dst.raw[:n]

    def compress_text(self, text: str) -> bytes:
        return self.compress(text.encode("utf-8"))

    def decompress_text(self, data: bytes, *, max_output_size: int = DEFAULT_MAX_OUTPUT_SIZE) -> str:
        return self.decompress(data, max_output_size=max_output_size).decode("utf-8")


def train_dictionary(samples: list[bytes], *, dict_size: int = 112_640) -> bytes:
    """Train a zstd dictionary (ZDICT) from sample documents."""
    if _LIB is None or not hasattr(_LIB, "ZDICT_trainFromBuffer"):
        raise ZstdError("dictionary training needs libzstd")
    blob = b"".join(samples)
    sizes = (ctypes.c_size_t * len(samples))(*[len(s) for s in samples])
    dst = ctypes.create_string_buffer(dict_size)
    n = _LIB.ZDICT_trainFromBuffer(dst, dict_size, blob, sizes, len(samples))
    if _LIB.ZSTD_isError(n):
        raise ZstdError(_LIB.ZSTD_getErrorName(n).decode())
    return dst.raw[:n]
