"""Stand-in for the `zstandard` wheel (absent from this image) so the UNMODIFIED reference can be imported.

Uses the system libzstd through ctypes when it is present (real zstd frames, so data written here is readable by a real
install), otherwise zlib.  The reference benchmark arm runs with compression disabled (the reference's default)."""
from __future__ import annotations

import ctypes
import ctypes.util
import zlib


class ZstdError(Exception):
    pass


def _load():
    for name in ("libzstd.so.1", ctypes.util.find_library("zstd")):
        if not name:
            continue
        try:
            lib = ctypes.CDLL(name)
        except OSError:
            continue
        lib.ZSTD_compressBound.restype = ctypes.c_size_t
        lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        lib.ZSTD_compress.restype = ctypes.c_size_t
        lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
        lib.ZSTD_decompress.restype = ctypes.c_size_t
        lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
        lib.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
        lib.ZSTD_getFrameContentSize.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        lib.ZSTD_isError.restype = ctypes.c_uint
        lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
        return lib
    return None


_LIB = _load()
_ZLIB_TAG = b"\x00zl1"


class ZstdCompressionDict:
    def __init__(self, data=b"", *a, **k):
        self._data = bytes(data)

    def as_bytes(self):
        return self._data


class ZstdCompressor:
    def __init__(self, level=3, dict_data=None, **_k):
        self.level = int(level)

    def compress(self, data):
        data = bytes(data)
        if _LIB is None:
            return _ZLIB_TAG + zlib.compress(data, min(9, max(1, self.level // 2)))
        cap = _LIB.ZSTD_compressBound(len(data))
        buf = ctypes.create_string_buffer(cap)
        n = _LIB.ZSTD_compress(buf, cap, data, len(data), self.level)
        if _LIB.ZSTD_isError(n):
            raise ZstdError("compress failed")
        return buf.raw[:n]


class ZstdDecompressor:
    def __init__(self, dict_data=None, **_k):
        pass

    def decompress(self, data, max_output_size=0):
        data = bytes(data)
        if data.startswith(_ZLIB_TAG):
            return zlib.decompress(data[len(_ZLIB_TAG):])
        if _LIB is None:
            raise ZstdError("libzstd unavailable")
        size = _LIB.ZSTD_getFrameContentSize(data, len(data))
        if size in (2 ** 64 - 1, 2 ** 64 - 2):
            size = max_output_size or (len(data) * 64 + 1024)
        if max_output_size and size > max_output_size:
            raise ZstdError("decompressed size exceeds limit")
        buf = ctypes.create_string_buffer(int(size) or 1)
        n = _LIB.ZSTD_decompress(buf, int(size), data, len(data))
        if _LIB.ZSTD_isError(n):
            raise ZstdError("decompress failed")
        return buf.raw[:n]


def train_dictionary(dict_size, samples, *a, **k):
    return ZstdCompressionDict(b"".join(samples)[:dict_size])
