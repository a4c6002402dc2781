"""Attention kernels (csrc/attn/attn_fwd.cu): tcgen05 flash-style forward + single-token decode.

K7 of SURVEY.md §2.4.  ``attention_ref`` is the fp32 PyTorch oracle used by the tests.
"""
from __future__ import annotations

import ctypes
import math

import torch

from infomesh_b200 import _native

LOG2E = 1.4426950408889634


def attention_ref(q, k, v, n_heads, kv_lens=None, causal=False, causal_offset=0, scale=None, rel_bias=None):
    """q: [B, Sq, nH*hd], k/v: [B, Sk, nH*hd] -> [B, Sq, nH*hd] (fp32).

    ``rel_bias``: [nH, Sq+Sk-1] natural-log additive bias indexed by (j - i) + (Sq - 1).
    """
    B, Sq, HH = q.shape
    Sk = k.shape[1]
    hd = HH // n_heads
    scale = (1.0 / math.sqrt(hd)) if scale is None else scale
    qf = q.float().view(B, Sq, n_heads, hd).transpose(1, 2)
    kf = k.float().view(B, Sk, n_heads, hd).transpose(1, 2)
    vf = v.float().view(B, Sk, n_heads, hd).transpose(1, 2)
    s = (qf @ kf.transpose(-1, -2)) * scale
    i = torch.arange(Sq, device=q.device)[:, None]
    j = torch.arange(Sk, device=q.device)[None, :]
    if rel_bias is not None:
        s = s + rel_bias.float()[:, (j - i) + (Sq - 1)][None]
    mask = torch.ones(B, 1, Sq, Sk, dtype=torch.bool, device=q.device)
    if kv_lens is not None:
        mask = mask & (j[None, None] < kv_lens.view(B, 1, 1, 1))
    if causal:
        mask = mask & (j <= i + causal_offset)[None, None]
    s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)
    o = p @ vf
    return o.transpose(1, 2).reshape(B, Sq, HH)


def attention(q, k, v, n_heads, kv_lens=None, causal=False, causal_offset=0, scale=None, rel_bias=None, out=None,
              cu_seqlens=None):
    """Fused attention.  ``q``: [B, Sq, nH*hd] view (may be a column slice of a packed QKV buffer).

    ``cu_seqlens`` (int32 ``[B+1]``): the buffers hold an UNPADDED batch -- sequence b owns rows
    ``[cu[b], cu[b+1])`` of the flattened ``[B*Sq, .]`` storage (``Sq`` = max sequence length <= 128)."""
    assert q.is_cuda and q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16
    B, Sq, HH = q.shape
    Sk = k.shape[1]
    hd = HH // n_heads
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    assert q.stride(0) == Sq * q.stride(1) and k.stride(0) == Sk * k.stride(1) and v.stride(0) == Sk * v.stride(1)
    scale = (1.0 / math.sqrt(hd)) if scale is None else scale
    if out is None:
        out = torch.empty((B, Sq, HH), device=q.device, dtype=torch.bfloat16)
    assert out.stride(2) == 1 and out.stride(0) == Sq * out.stride(1)
    bias_dev = None
    if rel_bias is not None:
        bias_dev = (rel_bias.float() * LOG2E).contiguous()
        assert bias_dev.shape == (n_heads, Sq + Sk - 1)
    if kv_lens is not None:
        assert kv_lens.dtype == torch.int32 and kv_lens.numel() == B
    L = _native.require()
    rc = L.im_attn_fwd(_native.ptr(q), _native.ptr(k), _native.ptr(v), _native.ptr(out), ctypes.c_int(B),
                       ctypes.c_int(n_heads), ctypes.c_int(hd), ctypes.c_int(Sq), ctypes.c_int(Sk),
                       ctypes.c_int(q.stride(1)), ctypes.c_int(k.stride(1)), ctypes.c_int(v.stride(1)),
                       ctypes.c_int(out.stride(1)), _native.ptr(kv_lens), ctypes.c_int(1 if causal else 0),
                       ctypes.c_int(causal_offset), ctypes.c_float(scale), _native.ptr(bias_dev),
                       _native.stream_ptr(), _native.ptr(cu_seqlens))
    _native.check(rc, "im_attn_fwd")
    _native.count_launch()
    return out


def attention_mx(q, k, v, n_heads, mx_out, kv_lens=None, scale=None, cu_seqlens=None):
    """Single-chunk (S <= 128, head_dim 64) attention whose context is written as MXFP8 into ``mx_out`` (an
    :class:`infomesh_b200.ops.mx.MxTensor` over the flattened ``[B*S, nH*hd]`` rows): the quantiser of the
    out-projection's input is the attention epilogue."""
    assert q.is_cuda and q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16
    B, Sq, HH = q.shape
    Sk = k.shape[1]
    hd = HH // n_heads
    assert hd == 64 and Sq <= 128 and Sk <= 128 and mx_out.q.shape == (B * Sq, HH)
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    scale = (1.0 / math.sqrt(hd)) if scale is None else scale
    L = _native.require()
    rc = L.im_attn_fwd_mx(_native.ptr(q), _native.ptr(k), _native.ptr(v), _native.ptr(mx_out.q),
                          ctypes.c_int(mx_out.q.stride(0)), _native.ptr(mx_out.sf), ctypes.c_int(HH // 128), ctypes.c_int(B),
                          ctypes.c_int(n_heads), ctypes.c_int(hd), ctypes.c_int(Sq), ctypes.c_int(Sk),
                          ctypes.c_int(q.stride(1)), ctypes.c_int(k.stride(1)), ctypes.c_int(v.stride(1)),
                          _native.ptr(kv_lens), ctypes.c_float(scale), _native.stream_ptr(), _native.ptr(cu_seqlens))
    _native.check(rc, "im_attn_fwd_mx")
    _native.count_launch()
    return mx_out


def attention_decode(q, k_cache, v_cache, n_heads, kv_len, scale=None, rel_bias_log2=None, q_pos=0, out=None,
                     seq_start=None, step_dev=None):
    """One query token per sequence against a KV cache ``[B, S_max, nH*hd]``; ``kv_len`` int or int32 tensor.

    ``seq_start`` (int32 ``[B]``): the caches are one packed ``[1, rows, nH*hd]`` buffer and sequence b's keys start
    at row ``seq_start[b]`` (CLS-only last layer of an unpadded cross-encoder).  ``step_dev`` (int32 ``[1]``): decode
    step ``t`` read on the device -- keys ``[0, t]``, query position ``t`` -- so one decoder step can replay from a
    CUDA graph."""
    B, HH = q.shape
    hd = HH // n_heads
    s_max = k_cache.shape[1]
    scale = (1.0 / math.sqrt(hd)) if scale is None else scale
    if out is None:
        out = torch.empty((B, HH), device=q.device, dtype=torch.bfloat16)
    lens_t = kv_len if torch.is_tensor(kv_len) else None
    bias_len = rel_bias_log2.shape[1] if rel_bias_log2 is not None else 0
    L = _native.require()
    rc = L.im_attn_decode(_native.ptr(q), ctypes.c_int(q.stride(0)), _native.ptr(k_cache), _native.ptr(v_cache),
                          ctypes.c_int(k_cache.stride(1)), ctypes.c_int(s_max), _native.ptr(lens_t),
                          ctypes.c_int(0 if lens_t is not None else int(kv_len)), ctypes.c_float(scale),
                          _native.ptr(rel_bias_log2), ctypes.c_int(bias_len), ctypes.c_int(q_pos), ctypes.c_int(B),
                          ctypes.c_int(n_heads), ctypes.c_int(hd), _native.ptr(out), ctypes.c_int(out.stride(0)),
                          _native.stream_ptr(), _native.ptr(seq_start), _native.ptr(step_dev))
    _native.check(rc, "im_attn_decode")
    _native.count_launch()
    return out


def kv_append(qkv, k_cache, v_cache, step_dev):
    """``k_cache[:, t] / v_cache[:, t] <- qkv[:, inner:2*inner] / qkv[:, 2*inner:]`` with ``t = step_dev[0]`` read on the
    device.  ``qkv``: bf16 ``[B, 3*inner]``; caches: contiguous ``[B, S_max, inner]``."""
    B, three = qkv.shape
    inner = three // 3
    assert k_cache.is_contiguous() and v_cache.is_contiguous() and k_cache.shape == (B, k_cache.shape[1], inner)
    L = _native.require()
    rc = L.im_kv_append(_native.ptr(qkv), ctypes.c_int(qkv.stride(0)), ctypes.c_int(inner), _native.ptr(k_cache),
                        _native.ptr(v_cache), ctypes.c_int(k_cache.shape[1]), ctypes.c_int(B), _native.ptr(step_dev),
                        _native.stream_ptr())
    _native.check(rc, "im_kv_append")
    _native.count_launch()
