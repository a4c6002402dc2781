"""tcgen05 GEMM (csrc/gemm/gemm_bf16.cu): ``C = act(alpha * A @ B^T + bias) + residual``.

K6/K8 of SURVEY.md §2.4 — the linear layers the reference runs through sentence-transformers
(reference infomesh/index/vector_store.py:104-125).  ``linear_ref`` is the fp32 oracle.
"""
from __future__ import annotations

import ctypes

import torch

from infomesh_b200 import _native

ACT = {None: 0, "none": 0, "gelu": 1, "relu": 2, "gelu_tanh": 3, "tanh": 4}


def linear_ref(a, w, bias=None, residual=None, act=None, alpha=1.0):
    """fp32 PyTorch reference of :func:`linear`."""
    y = alpha * (a.float() @ w.float().t())
    if bias is not None:
        y = y + bias.float()
    if act in ("gelu", 1):
        y = torch.nn.functional.gelu(y)
    elif act in ("relu", 2):
        y = torch.relu(y)
    elif act in ("gelu_tanh", 3):
        y = torch.nn.functional.gelu(y, approximate="tanh")
    elif act in ("tanh", 4):
        y = torch.tanh(y)
    if residual is not None:
        y = y + residual.float()
    return y


_FP8_DTYPES = (torch.uint8, torch.float8_e4m3fn)


def linear(a, w, bias=None, residual=None, act=None, alpha=1.0, out=None, out_dtype=torch.bfloat16, bn=0,
           max_ctas=0, rs=None, ag=None, m_dev=None, row_scale=None):
    """``a[M,K] @ w[N,K]^T`` with fused bias / activation / residual on the tcgen05 kernel.

    ``m_dev``: optional int32 device scalar with the number of valid rows (<= M) -- the kernel reads it at run time,
    so unpadded batches whose token count changes per step still replay from one CUDA graph.

    fp8: pass ``a`` / ``w`` as e4m3 bytes (``torch.float8_e4m3fn`` or ``uint8``, see :func:`quantize_rows_fp8` /
    :func:`quantize_weight_fp8`); ``row_scale`` (fp32 ``[M]``) and ``alpha`` (= weight scale) dequantise in the
    epilogue.  The MMAs run as ``kind::f8f6f4`` with fp32 accumulation; output, bias, activation, residual as usual.

    Tensor-parallel hooks (``parallel.tp``): ``rs`` = :class:`ReduceScatterChannel` — the epilogue pushes every
    128-row block of the partial product into the owning rank's receive slot over NVLink instead of storing locally;
    ``ag`` = :class:`AllGatherInput` — ``a`` is a full-sequence buffer that peers are still filling, the TMA producer
    waits per row block on the arrival counters and starts with this rank's own rows."""
    fp8 = a.dtype in _FP8_DTYPES
    assert a.is_cuda and w.is_cuda and ((fp8 and w.dtype in _FP8_DTYPES) or (a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16))
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    assert a.stride(1) == 1 and w.stride(1) == 1
    m, k = a.shape
    n = w.shape[0]
    if out is None and rs is None:
        out = torch.empty((m, n), device=a.device, dtype=out_dtype)
    if rs is not None:
        assert m == rs.rows_per_rank * rs.world and n == rs.n_cols and residual is None
        out = rs.recv[rs.rank]                     # placeholder view: the kernel addresses peers through rs.peer_c
        out = out.view(rs.rows_per_rank, n)
    assert out.stride(1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.shape == (m, n) and residual.stride(1) == 1
    L = _native.require()
    rc = L.im_gemm_bf16_tn(
        _native.ptr(a), _native.ptr(w), _native.ptr(out), _native.ptr(bias), _native.ptr(residual),
        ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k),
        ctypes.c_int(a.stride(0)), ctypes.c_int(w.stride(0)), ctypes.c_int(out.stride(0)),
        ctypes.c_int(residual.stride(0) if residual is not None else 0),
        ctypes.c_int(ACT[act] if not isinstance(act, int) else act),
        ctypes.c_int(1 if out.dtype == torch.float32 else 0), ctypes.c_float(alpha), ctypes.c_int(bn),
        ctypes.c_void_p(rs.peer_c_ptr if rs else 0), ctypes.c_void_p(rs.peer_flags_ptr if rs else 0),
        ctypes.c_int(rs.rank if rs else 0), ctypes.c_int(rs.rows_per_rank if rs else 0),
        ctypes.c_void_p(ag.flags_ptr if ag else 0), ctypes.c_void_p(ag.state_ptr if ag else 0),
        ctypes.c_int(ag.m_rotate if ag else 0), ctypes.c_int(max_ctas or _native.sm_budget("gemm")), _native.stream_ptr(),
        rs.peer_c_host if rs else ctypes.c_void_p(0), ctypes.c_int(rs.world if rs else 0),
        _native.ptr(m_dev), ctypes.c_int(1 if fp8 else 0), _native.ptr(row_scale))
    _native.check(rc, "im_gemm_bf16_tn")
    _native.count_launch()
    return out


def quantize_rows_fp8(x, n_rows_dev=None):
    """bf16 ``[M, K]`` -> (e4m3 bytes ``[M, K]``, fp32 per-row scales ``[M]``): dynamic per-token scaling."""
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    m, k = x.shape
    q = torch.empty((m, k), device=x.device, dtype=torch.uint8)
    scale = torch.ones((m,), device=x.device, dtype=torch.float32)
    L = _native.require()
    rc = L.im_quantize_rows_fp8(_native.ptr(x), ctypes.c_int(x.stride(0)), ctypes.c_int(k), ctypes.c_int(m), _native.ptr(q),
                                ctypes.c_int(q.stride(0)), _native.ptr(scale), _native.ptr(n_rows_dev), _native.stream_ptr())
    _native.check(rc, "im_quantize_rows_fp8")
    _native.count_launch()
    return q, scale


def quantize_weight_fp8(w):
    """bf16 weight ``[N, K]`` -> (e4m3 bytes, python-float per-tensor scale).  One-off at load time (plain torch)."""
    amax = float(w.float().abs().max().item())
    scale = amax / 448.0 if amax > 0 else 1.0
    w8 = (w.float() / scale).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8).contiguous()
    return w8, scale


def linear_fp8_ref(a8, a_scale, w8, w_scale, bias=None, residual=None, act=None):
    """fp32 oracle of the fp8 path: dequantise, then :func:`linear_ref`."""
    a = a8.view(torch.float8_e4m3fn).float() * a_scale[:, None]
    w = w8.view(torch.float8_e4m3fn).float() * w_scale
    return linear_ref(a, w, bias, residual, act)

