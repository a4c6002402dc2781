"""ctypes bridge to ``libinfomesh_b200.so`` (CUDA kernels for sm_100a + the C++ host runtime).

The library is built in-tree by :mod:`infomesh_b200.build`.  Kernels are launched on the
*current torch CUDA stream* by passing ``tensor.data_ptr()`` and the raw stream handle, so they
compose with CUDA-graph capture and torch's stream semantics without a torch C++ extension.

Importing this module never requires a GPU.  ``lib()`` raises :class:`NativeUnavailable` when
the shared object is missing and cannot be built; GPU ops call ``require()`` so a missing
extension on a GPU box fails loudly instead of silently falling back.
"""
from __future__ import annotations

import ctypes
import os
import threading
from pathlib import Path

# INFOMESH_B200_NATIVE_LIB points the bridge at another build of the library (the sanitizer build of the host runtime)
_LIB_PATH = Path(os.environ.get("INFOMESH_B200_NATIVE_LIB") or Path(__file__).resolve().parent / "_native" / "libinfomesh_b200.so")
_lock = threading.Lock()
_lib: ctypes.CDLL | None = None
_load_error: str | None = None


class NativeUnavailable(RuntimeError):
    """The native library could not be loaded (not built, or nvcc missing)."""


class NativeError(RuntimeError):
    """A native call returned a non-zero status."""


def lib_path() -> Path:
    return _LIB_PATH


def lib(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building on first use when possible) and return the shared library."""
    global _lib, _load_error
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not _LIB_PATH.exists() and build_if_missing and os.environ.get("INFOMESH_B200_NO_BUILD") != "1":
            try:
                from infomesh_b200.build import build

                build()
            except Exception as exc:  # noqa: BLE001 - surfaced through NativeUnavailable
                _load_error = f"build failed: {exc}"
        if not _LIB_PATH.exists():
            raise NativeUnavailable(_load_error or f"{_LIB_PATH} not built (run python -m infomesh_b200.build)")
        try:
            handle = ctypes.CDLL(str(_LIB_PATH))
        except OSError as exc:
            _load_error = str(exc)
            raise NativeUnavailable(_load_error) from exc
        if hasattr(handle, "im_last_error"):          # absent from host-only builds
            handle.im_last_error.restype = ctypes.c_char_p
        _lib = handle
        return handle


def available() -> bool:
    try:
        lib()
        return True
    except NativeUnavailable:
        return False


def require() -> ctypes.CDLL:
    """Like :func:`lib` but with an explicit message for GPU code paths."""
    try:
        return lib()
    except NativeUnavailable as exc:
        raise NativeUnavailable(
            f"infomesh_b200 native kernels are required on this path and are not loadable: {exc}"
        ) from exc


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().im_last_error()
        raise NativeError(f"{what} failed: {msg.decode(errors='replace') if msg else rc}")


def ptr(t) -> ctypes.c_void_p:
    """Device/host pointer of a torch tensor (``None`` -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr() -> ctypes.c_void_p:
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_launch_count = 0


def count_launch(n: int = 1) -> None:
    """Book-keeping of native kernel launches (reported by bench.py as ``gpu_launches``)."""
    global _launch_count
    _launch_count += n


def launch_count() -> int:
    return _launch_count


def reset_launch_count() -> None:
    global _launch_count
    _launch_count = 0


# ---------------------------------------------------------------------------------------------------------------------
# SM budgets.  The persistent kernels (GEMMs, similarity scan) launch one CTA per SM by default.  The serving pipeline
# overlaps an HBM-bound scan with tensor-core-bound GEMMs on two streams; both kinds need a whole SM's shared memory, so
# they only run side by side if each is told to leave SMs free.  ``sm_budget(kind)`` is what a wrapper uses when its
# caller passed ``max_ctas=0``; ``with sm_budgets(gemm=136, scan=12): ...`` scopes a setting (launch dimensions are frozen
# into a CUDA graph at capture time, so the scope only has to cover warm-up + capture).
SM_BUDGET = {"gemm": 0, "scan": 0}


def sm_budget(kind: str) -> int:
    return SM_BUDGET.get(kind, 0)


class sm_budgets:
    def __init__(self, **limits: int):
        self._new, self._old = limits, {}

    def __enter__(self):
        for kind, n in self._new.items():
            self._old[kind] = SM_BUDGET.get(kind, 0)
            SM_BUDGET[kind] = int(n)
        return self

    def __exit__(self, *exc):
        SM_BUDGET.update(self._old)
        return False
