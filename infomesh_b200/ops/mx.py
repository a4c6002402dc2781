"""MXFP8 (block-scaled e4m3) tensors and the tcgen05 ``kind::mxf8f6f4.block_scale`` GEMM (csrc/gemm/gemm_mxf8.cu).

An MX tensor is ``q`` (e4m3 bytes ``[rows, K]``) plus one ue8m0 scale per row and 32 K-elements.  On the device the
scales live in 512-byte chunks already shaped for ``tcgen05.cp.32x128b.warpx4`` (128 rows x 4 scales per chunk, byte
``(r % 32) * 16 + (r // 32) * 4 + s``):

* activations / GEMM "A" side: ``[ceil(M/128)][K/128]`` chunks — written directly by the producers (LayerNorm,
  attention, the GELU epilogue of the previous GEMM), never by a standalone quantiser on the hot path;
* weights / "B" side: ``[K/128][n_chunks]`` chunks (k-block major), packed once at load time.

The plain-torch functions here are the oracles for those kernels and the load-time weight packer.
Replaces the fp32 ``model.encode`` GEMMs of the reference (infomesh/index/vector_store.py:104-125) and the LLM-over-HTTP
reranker (infomesh/search/reranker.py:124-159) with block-scaled fp8 tensor-core math."""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import torch

from infomesh_b200 import _native
from infomesh_b200.ops.gemm import ACT, linear_ref

E4M3_MAX = 448.0


# ----------------------------------------------------------------------------- torch oracles / packers
def quantize_ref(x: torch.Tensor):
    """fp32/bf16 ``[R, K]`` (K % 32 == 0) -> (e4m3 bytes ``[R, K]`` uint8, ue8m0 exponents ``[R, K/32]`` uint8).

    Scale = smallest power of two with ``amax / scale <= 448`` (exponent clamped to [1, 253]) -- bit-identical to
    ``ue8m0_from_amax`` in csrc/common/ptx.cuh."""
    R, K = x.shape
    xf = x.float().reshape(R, K // 32, 32)
    amax = xf.abs().amax(-1)
    t = (amax * (1.0 / E4M3_MAX)).contiguous()
    e = ((t.view(torch.int32) + 0x7FFFFF) >> 23).clamp(1, 253)
    inv = ((254 - e) << 23).to(torch.int32).view(torch.float32)
    q = (xf * inv[..., None]).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).view(torch.uint8).reshape(R, K)
    return q.contiguous(), e.to(torch.uint8).contiguous()


def dequantize(q: torch.Tensor, e: torch.Tensor) -> torch.Tensor:
    """(e4m3 bytes ``[R, K]``, exponents ``[R, K/32]``) -> fp32 ``[R, K]``."""
    R, K = q.shape
    scale = (e.to(torch.int32) << 23).view(torch.float32)        # 2^(e - 127)
    return (q.view(torch.float8_e4m3fn).float().reshape(R, K // 32, 32) * scale[..., None]).reshape(R, K)


def _chunk_rows(e: torch.Tensor, n_row_blocks: int) -> torch.Tensor:
    """exponents ``[R, K/32]`` -> ``[n_row_blocks, K/128, 512]`` chunk bytes (rows padded with exponent 127 = 1.0)."""
    R, S = e.shape
    assert S % 4 == 0
    pad = n_row_blocks * 128 - R
    if pad:
        e = torch.cat([e, torch.full((pad, S), 127, dtype=torch.uint8, device=e.device)], 0)
    # [rb, m1(4), m0(32), kb, s(4)] -> [rb, kb, m0, m1, s]
    v = e.reshape(n_row_blocks, 4, 32, S // 4, 4).permute(0, 3, 2, 1, 4)
    return v.reshape(n_row_blocks, S // 4, 512).contiguous()


def pack_sfa(e: torch.Tensor) -> torch.Tensor:
    """A-side layout ``[ceil(M/128), K/128, 512]``."""
    return _chunk_rows(e, (e.shape[0] + 127) // 128)


def sfb_chunks(n: int) -> int:
    """chunks per k-block the kernel may touch for an N-column weight (192-row tiles straddle two 128-row chunks)."""
    return ((n + 191) // 192 * 192 + 127) // 128


def pack_sfb(e: torch.Tensor) -> torch.Tensor:
    """B-side layout ``[K/128, n_chunks, 512]`` (k-block major)."""
    nc = sfb_chunks(e.shape[0])
    return _chunk_rows(e, nc).permute(1, 0, 2).contiguous()


def unpack_sfa(chunks: torch.Tensor, rows: int) -> torch.Tensor:
    """inverse of :func:`pack_sfa`: ``[rb, kb, 512]`` -> exponents ``[rows, 4 * kb]``."""
    rb, kb, _ = chunks.shape
    v = chunks.reshape(rb, kb, 32, 4, 4).permute(0, 3, 2, 1, 4)      # [rb, m1, m0, kb, s]
    return v.reshape(rb * 128, kb * 4)[:rows].contiguous()


@dataclass
class MxTensor:
    """An activation (or any GEMM A operand) in MXFP8: ``q`` uint8 ``[M, K]`` + A-side scale chunks."""
    q: torch.Tensor
    sf: torch.Tensor          # uint8 [ceil(M/128), K/128, 512]

    @property
    def shape(self):
        return self.q.shape

    def float(self, rows: int | None = None) -> torch.Tensor:
        r = self.q.shape[0] if rows is None else rows
        return dequantize(self.q[:r], unpack_sfa(self.sf, r))


@dataclass
class MxWeight:
    """A weight ``[N, K]`` in MXFP8 with B-side scale chunks (packed once at load time by :func:`quantize_weight`)."""
    q: torch.Tensor
    sf: torch.Tensor          # uint8 [K/128, n_chunks, 512]
    e: torch.Tensor           # exponents [N, K/32] (oracle / debugging)

    @property
    def n_chunks(self) -> int:
        return self.sf.shape[1]

    def float(self) -> torch.Tensor:
        return dequantize(self.q, self.e)


def quantize_weight(w: torch.Tensor) -> MxWeight:
    q, e = quantize_ref(w)
    return MxWeight(q, pack_sfb(e), e)


def quantize_act_ref(x: torch.Tensor) -> MxTensor:
    """torch build of an :class:`MxTensor` (tests, and cold paths that have no fused producer)."""
    q, e = quantize_ref(x)
    return MxTensor(q, pack_sfa(e))


def alloc_act(m: int, k: int, device, init: bool = False) -> MxTensor:
    """MX activation buffer for a fused producer to fill.  ``init=False`` (hot path): no fill kernels -- rows the
    producer skips (beyond a varlen batch's token count) hold garbage that only ever reaches their own, equally
    ignored, output rows.  ``init=True`` zeroes the data and sets every scale to 1.0 (tests, dequantising whole buffers)."""
    if not init:
        return MxTensor(torch.empty((m, k), device=device, dtype=torch.uint8),
                        torch.empty(((m + 127) // 128, k // 128, 512), device=device, dtype=torch.uint8))
    q = torch.zeros((m, k), device=device, dtype=torch.uint8)
    sf = torch.full(((m + 127) // 128, k // 128, 512), 127, device=device, dtype=torch.uint8)
    return MxTensor(q, sf)


def linear_mx_ref(a: MxTensor, w: MxWeight, bias=None, residual=None, act=None, rows=None):
    """fp32 oracle: dequantise both operands, then the plain reference linear."""
    return linear_ref(a.float(rows), w.float(), bias, residual, act)


# ----------------------------------------------------------------------------- kernel
def linear_mx(a: MxTensor, w: MxWeight, bias=None, residual=None, act=None, out=None, out_mx: bool = False, m_dev=None,
              max_ctas: int = 0, mode: int = 0):
    """``a @ w^T`` on ``tcgen05.mma.kind::mxf8f6f4.block_scale`` with fused bias / activation / residual.

    ``out_mx=False`` -> bf16 ``[M, N]``;  ``out_mx=True`` -> :class:`MxTensor` ``[M, N]`` quantised in the epilogue (the next
    GEMM's A operand).  ``mode``: 0 auto, 1 operand-ring kernel, 2 A-resident kernel (K <= 768)."""
    m, k = a.q.shape
    n = w.q.shape[0]
    assert a.q.is_cuda and a.q.dtype == torch.uint8 and w.q.dtype == torch.uint8 and w.q.shape[1] == k and k % 128 == 0
    assert a.q.stride(1) == 1 and w.q.stride(1) == 1 and a.sf.is_contiguous() and w.sf.is_contiguous()
    assert a.sf.shape[0] >= (m + 127) // 128 and a.sf.shape[1] == k // 128
    dev = a.q.device
    if out_mx:
        assert residual is None and n % 128 == 0
        if out is None:
            out = alloc_act(m, n, dev)
        c, c_sf, ldc = out.q, out.sf, out.q.stride(0)
    else:
        if out is None:
            out = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
        assert out.dtype == torch.bfloat16 and out.stride(1) == 1
        c, c_sf, ldc = out, None, out.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.shape == (m, n) and residual.stride(1) == 1
    L = _native.require()
    rc = L.im_gemm_mxf8(_native.ptr(a.q), _native.ptr(a.sf), _native.ptr(w.q), _native.ptr(w.sf), ctypes.c_int(w.n_chunks),
                        _native.ptr(c), _native.ptr(c_sf), _native.ptr(bias), _native.ptr(residual),
                        ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k), ctypes.c_int(a.q.stride(0)),
                        ctypes.c_int(w.q.stride(0)), ctypes.c_int(ldc),
                        ctypes.c_int(residual.stride(0) if residual is not None else 0),
                        ctypes.c_int(ACT[act] if not isinstance(act, int) else act), ctypes.c_int(1 if out_mx else 0),
                        _native.ptr(m_dev), ctypes.c_int(max_ctas or _native.sm_budget("gemm")), _native.stream_ptr(), ctypes.c_int(mode))
    if rc < 0:
        _native.check(rc, "im_gemm_mxf8")
    _native.count_launch()
    return out


FUSED_LN = os.environ.get("INFOMESH_B200_FUSED_LN", "1") != "0"     # models/bert.py: LayerNorm in the projection's epilogue
FUSED_LN_WIDTHS = (384, 768)        # N = 2 or 4 tiles of 192 columns: a cluster of 2 / 4 CTAs owns a 128-row block


def linear_mx_ln(a: MxTensor, w: MxWeight, bias, residual, gamma, beta, eps: float, mx_out: MxTensor, out=None, m_dev=None,
                 max_ctas: int = 0):
    """``LN(a @ w^T + bias + residual)`` -> bf16 ``[M, N]`` AND its MXFP8 copy in ``mx_out``, in one kernel.

    The attention-output and FFN-down projections of a post-LN transformer block end in residual add + LayerNorm; with
    ``N`` = 384 / 768 the ``N / 192`` CTAs that compute one 128-row block run as a thread-block cluster, exchange per-row
    (sum, sum of squares) through distributed shared memory and normalise in the GEMM epilogue, so the ``[M, N]``
    activations never make the extra HBM round trip through a standalone LayerNorm kernel."""
    m, k = a.q.shape
    n = w.q.shape[0]
    assert n in FUSED_LN_WIDTHS, f"fused LayerNorm epilogue needs N in {FUSED_LN_WIDTHS}"
    assert a.q.is_cuda and a.q.dtype == torch.uint8 and w.q.dtype == torch.uint8 and w.q.shape[1] == k and k % 128 == 0
    assert a.q.stride(1) == 1 and w.q.stride(1) == 1 and a.sf.is_contiguous() and w.sf.is_contiguous()
    assert mx_out.q.shape == (m, n) and mx_out.q.stride(1) == 1 and mx_out.sf.is_contiguous()
    assert gamma.dtype == torch.float32 and gamma.numel() == n and (beta is None or (beta.dtype == torch.float32 and beta.numel() == n))
    assert bias is None or (bias.dtype == torch.float32 and bias.numel() == n)
    if out is None:
        out = torch.empty((m, n), device=a.q.device, dtype=torch.bfloat16)
    assert out.dtype == torch.bfloat16 and out.stride(1) == 1
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.shape == (m, n) and residual.stride(1) == 1
    L = _native.require()
    rc = L.im_gemm_mxf8_ln(_native.ptr(a.q), _native.ptr(a.sf), _native.ptr(w.q), _native.ptr(w.sf), ctypes.c_int(w.n_chunks),
                           _native.ptr(out), _native.ptr(mx_out.q), _native.ptr(mx_out.sf), _native.ptr(bias), _native.ptr(residual),
                           _native.ptr(gamma), _native.ptr(beta), ctypes.c_float(eps),
                           ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k), ctypes.c_int(a.q.stride(0)), ctypes.c_int(w.q.stride(0)),
                           ctypes.c_int(out.stride(0)), ctypes.c_int(residual.stride(0) if residual is not None else 0),
                           ctypes.c_int(mx_out.q.stride(0)), _native.ptr(m_dev), ctypes.c_int(max_ctas or _native.sm_budget("gemm")),
                           _native.stream_ptr())
    if rc < 0:
        _native.check(rc, "im_gemm_mxf8_ln")
    _native.count_launch()
    return out


def linear_mx_ln_ref(a: MxTensor, w: MxWeight, bias, residual, gamma, beta, eps: float, rows=None):
    """fp32 oracle of :func:`linear_mx_ln`: dequantised GEMM + bias + residual, then LayerNorm."""
    y = linear_mx_ref(a, w, bias, residual, None, rows)
    return torch.nn.functional.layer_norm(y, (y.shape[-1],), gamma.float(), None if beta is None else beta.float(), eps)
