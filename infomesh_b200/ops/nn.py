"""Transformer glue kernels (csrc/nn/norm_embed.cu) and their fp32 PyTorch oracles (K5/K9)."""
from __future__ import annotations

import ctypes

import torch

from infomesh_b200 import _native


# ----------------------------------------------------------------------------- references
def embed_ln_ref(ids, word, pos, type_emb, gamma, beta, eps, seq_len, pos_ids=None, type_ids=None, pos_offset=0):
    n = ids.numel()
    x = word.float()[ids.long().view(-1)]
    if pos is not None:
        p = pos_ids.long().view(-1) if pos_ids is not None else (torch.arange(n, device=ids.device) % seq_len) + pos_offset
        x = x + pos.float()[p]
    if type_emb is not None:
        t = type_ids.long().view(-1) if type_ids is not None else torch.zeros(n, dtype=torch.long, device=ids.device)
        x = x + type_emb.float()[t]
    if gamma is None:
        return x
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), gamma.float(), beta.float() if beta is not None else None, eps)


def layernorm_ref(x, gamma, beta, eps, residual=None, rms_only=False):
    x = x.float()
    if residual is not None:
        x = x + residual.float()
    if rms_only:
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * gamma.float()
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), gamma.float(), beta.float() if beta is not None else None, eps)


def pool_norm_ref(h, lengths, mode="cls", normalize=True):
    hf = h.float()
    if mode == "cls":
        e = hf[:, 0]
    else:
        B, S, _ = hf.shape
        m = (torch.arange(S, device=h.device)[None] < lengths.view(B, 1)).float()[..., None]
        e = (hf * m).sum(1) / m.sum(1).clamp(min=1)
    return torch.nn.functional.normalize(e, dim=-1) if normalize else e


# ----------------------------------------------------------------------------- kernels
def embed_ln(ids, word, pos, type_emb, gamma, beta, eps, seq_len, pos_ids=None, type_ids=None, pos_offset=0, out=None,
             n_rows_dev=None, mx_out=None):
    """``mx_out``: an :class:`infomesh_b200.ops.mx.MxTensor` that also receives the normalised rows as MXFP8."""
    n = ids.numel()
    H = word.shape[1]
    assert ids.dtype == torch.int32 and word.dtype == torch.bfloat16
    if out is None:
        out = torch.empty((n, H), device=word.device, dtype=torch.bfloat16)
    L = _native.require()
    if mx_out is not None:
        assert mx_out.q.shape == (n, H) and mx_out.sf.shape[1] == H // 128
        rc = L.im_embed_ln_mx(_native.ptr(ids), _native.ptr(pos_ids), _native.ptr(type_ids), _native.ptr(word),
                              _native.ptr(pos), _native.ptr(type_emb), _native.ptr(gamma), _native.ptr(beta),
                              ctypes.c_float(eps), ctypes.c_int(n), ctypes.c_int(seq_len), ctypes.c_int(pos_offset),
                              ctypes.c_int(word.shape[0]), ctypes.c_int(pos.shape[0] if pos is not None else 1),
                              ctypes.c_int(H), _native.ptr(out), _native.stream_ptr(), _native.ptr(n_rows_dev),
                              _native.ptr(mx_out.q), ctypes.c_int(mx_out.q.stride(0)), _native.ptr(mx_out.sf))
        _native.check(rc, "im_embed_ln_mx")
        _native.count_launch()
        return out
    rc = L.im_embed_ln(_native.ptr(ids), _native.ptr(pos_ids), _native.ptr(type_ids), _native.ptr(word),
                       _native.ptr(pos), _native.ptr(type_emb), _native.ptr(gamma), _native.ptr(beta),
                       ctypes.c_float(eps), ctypes.c_int(n), ctypes.c_int(seq_len), ctypes.c_int(pos_offset),
                       ctypes.c_int(word.shape[0]), ctypes.c_int(pos.shape[0] if pos is not None else 1),
                       ctypes.c_int(H), _native.ptr(out), _native.stream_ptr(), _native.ptr(n_rows_dev))
    _native.check(rc, "im_embed_ln")
    _native.count_launch()
    return out


def layernorm_mx(x, gamma, beta, eps, mx_out, residual=None, rms_only=False, out=None, want_bf16=True, n_rows_dev=None):
    """``LN(x + residual)`` written twice by ONE kernel: bf16 (the residual stream) and MXFP8 into ``mx_out`` (the A
    operand of the next block-scaled GEMM) -- the quantiser is fused into its producer."""
    n, H = x.shape
    assert x.dtype == torch.bfloat16 and x.stride(1) == 1 and mx_out.q.shape == (n, H)
    if out is None and want_bf16:
        out = torch.empty((n, H), device=x.device, dtype=torch.bfloat16)
    L = _native.require()
    rc = L.im_sum_ln_mx(_native.ptr(x), _native.ptr(residual), _native.ptr(gamma), _native.ptr(beta), ctypes.c_float(eps),
                        ctypes.c_int(1 if rms_only else 0), ctypes.c_int(n), ctypes.c_int(H),
                        _native.ptr(out if want_bf16 else None), _native.stream_ptr(), _native.ptr(n_rows_dev),
                        _native.ptr(mx_out.q), ctypes.c_int(mx_out.q.stride(0)), _native.ptr(mx_out.sf))
    _native.check(rc, "im_sum_ln_mx")
    _native.count_launch()
    return out


def layernorm(x, gamma, beta=None, eps=1e-12, residual=None, rms_only=False, out=None, sum_out=None, partials=1,
              partial_stride=0, want_norm=True, rs=None, ag_push=None, n_rows_dev=None):
    """``out = LN(sum_p x[p] + residual)``; ``x``: [n, H] (or [P, n, H] with ``partials=P``).

    ``rs``: consume a fused reduce-scatter (wait on the arrival counters of ``rs`` before summing its receive
    slots); ``ag_push``: also store the normalised rows into every peer's full-sequence buffer and bump their
    per-row-block counters (fused all-gather producer)."""
    H = x.shape[-1]
    n = x.shape[-2]
    assert x.dtype == torch.bfloat16 and x.stride(-1) == 1
    if out is None and want_norm:
        out = torch.empty((n, H), device=x.device, dtype=torch.bfloat16)
    L = _native.require()
    rc = L.im_sum_ln(_native.ptr(x), ctypes.c_longlong(partial_stride if partials > 1 else 0), ctypes.c_int(partials),
                     _native.ptr(residual), _native.ptr(gamma), _native.ptr(beta), ctypes.c_float(eps),
                     ctypes.c_int(1 if rms_only else 0), ctypes.c_int(n), ctypes.c_int(H),
                     _native.ptr(out if want_norm else None), _native.ptr(sum_out),
                     ctypes.c_void_p(rs.local_flags_ptr if rs else 0), ctypes.c_void_p(rs.step_ptr if rs else 0),
                     ctypes.c_uint(rs.arrivals_per_block if rs else 0), ctypes.c_int(rs.blocks_per_src if rs else 0),
                     ctypes.c_void_p(ag_push.peer_buf_ptr if ag_push else 0),
                     ctypes.c_void_p(ag_push.peer_flags_ptr if ag_push else 0),
                     ctypes.c_int(ag_push.row_offset if ag_push else 0), ctypes.c_int(ag_push.world if ag_push else 0),
                     ctypes.c_int(ag_push.rank if ag_push else 0), _native.stream_ptr(), _native.ptr(n_rows_dev))
    _native.check(rc, "im_sum_ln")
    _native.count_launch()
    return out


def seq_pack(ids, lens, pos_offset=0):
    """Padded ``ids[n, S]`` + ``lens[n]`` -> ``(packed_ids[n*S], packed_pos[n*S], cu_seqlens[n+1], total[1])``.

    Only the first ``total`` rows of the packed arrays are meaningful; ``total`` stays on the device (the GEMM / LN
    kernels read it at run time), so the unpadded forward replays from a CUDA graph whatever the batch's lengths."""
    n, S = ids.shape
    assert ids.dtype == torch.int32 and lens.dtype == torch.int32 and ids.is_contiguous()
    dev = ids.device
    packed_ids = torch.zeros((n * S,), device=dev, dtype=torch.int32)
    packed_pos = torch.zeros((n * S,), device=dev, dtype=torch.int32)
    cu = torch.empty((n + 1,), device=dev, dtype=torch.int32)
    total = torch.empty((1,), device=dev, dtype=torch.int32)
    L = _native.require()
    rc = L.im_seq_pack(_native.ptr(ids), _native.ptr(lens), ctypes.c_int(n), ctypes.c_int(S), ctypes.c_int(pos_offset),
                       _native.ptr(cu), _native.ptr(total), _native.ptr(packed_ids), _native.ptr(packed_pos),
                       _native.stream_ptr())
    _native.check(rc, "im_seq_pack")
    _native.count_launch()
    return packed_ids, packed_pos, cu, total


def gather_rows(x, idx, n):
    """``out[b] = x[idx[b]]`` for the first ``n`` entries of the int32 index vector (bf16 rows)."""
    assert x.dtype == torch.bfloat16 and x.stride(1) == 1 and idx.dtype == torch.int32
    H = x.shape[1]
    out = torch.empty((n, H), device=x.device, dtype=torch.bfloat16)
    L = _native.require()
    rc = L.im_gather_rows(_native.ptr(x), _native.ptr(idx), ctypes.c_int(n), ctypes.c_int(H), ctypes.c_int(x.stride(0)),
                          _native.ptr(out), _native.stream_ptr())
    _native.check(rc, "im_gather_rows")
    _native.count_launch()
    return out


def pool_norm(h, lengths=None, mode="cls", normalize=True, out=None, out_f32=None, out_q8=None, out_qscale=None):
    """Pooled (CLS / mean) and L2-normalised sentence vectors, bf16 ``[B, H]``.  ``out_q8`` (uint8 ``[B, H]``, e4m3 bytes)
    + ``out_qscale`` (fp32 ``[B]``, power of two): the same vectors quantised for the fp8 similarity search
    (``ops.search.sim_topk_f8``), emitted by the same kernel -- no separate quantisation pass."""
    B, S, H = h.shape
    assert h.is_contiguous() and h.dtype == torch.bfloat16
    if out is None:
        out = torch.empty((B, H), device=h.device, dtype=torch.bfloat16)
    if out_q8 is not None:
        assert out_q8.dtype == torch.uint8 and out_q8.is_contiguous() and out_q8.shape == (B, H)
        assert out_qscale is not None and out_qscale.dtype == torch.float32 and out_qscale.numel() >= B
    L = _native.require()
    rc = L.im_pool_norm(_native.ptr(h), _native.ptr(lengths), ctypes.c_int(B), ctypes.c_int(S), ctypes.c_int(H),
                        ctypes.c_int(0 if mode == "cls" else 1), ctypes.c_int(1 if normalize else 0),
                        _native.ptr(out), _native.ptr(out_f32), _native.stream_ptr(), _native.ptr(out_q8), _native.ptr(out_qscale))
    _native.check(rc, "im_pool_norm")
    _native.count_launch()
    return out


def quantize_rows_e4m3(x, normalize=False, chunk=1 << 20):
    """bf16 ``[n, H]`` -> (uint8 ``[n, H]`` e4m3 bytes, fp32 ``[n]`` power-of-two row scales) with the pooling kernel
    (sequence length 1): the index-build side of the fp8 dense shard.  ``value = e4m3 * scale``."""
    n, H = x.shape
    q8 = torch.empty((n, H), device=x.device, dtype=torch.uint8)
    sc = torch.empty((n,), device=x.device, dtype=torch.float32)
    scratch = torch.empty((min(n, chunk), H), device=x.device, dtype=torch.bfloat16)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        pool_norm(x[a:b].contiguous().view(b - a, 1, H), None, "cls", normalize, out=scratch[:b - a], out_q8=q8[a:b], out_qscale=sc[a:b])
    return q8, sc


def quantize_rows_e4m3_ref(x):
    """PyTorch oracle of :func:`quantize_rows_e4m3` (no normalisation): same power-of-two scale rule as the kernel."""
    xf = x.float()
    amax = xf.abs().amax(dim=1).clamp_min(1e-30)
    e = torch.ceil(torch.log2(amax / 448.0)).clamp(-126, 126)
    scale = torch.exp2(e)
    q = (xf / scale[:, None]).to(torch.float8_e4m3fn)
    return q.view(torch.uint8), scale


def cls_head(h, w1, b1, w2, b2, out=None):
    B, S, H = h.shape
    if out is None:
        out = torch.empty((B,), device=h.device, dtype=torch.float32)
    L = _native.require()
    rc = L.im_cls_head(_native.ptr(h), ctypes.c_int(B), ctypes.c_int(S), ctypes.c_int(H), _native.ptr(w1),
                       _native.ptr(b1), _native.ptr(w2), _native.ptr(b2), _native.ptr(out), _native.stream_ptr())
    _native.check(rc, "im_cls_head")
    _native.count_launch()
    return out


def row_argmax(x, id_offset=0, out_val=None, out_idx=None):
    n, c = x.shape
    assert x.dtype == torch.float32 and x.stride(1) == 1
    if out_val is None:
        out_val = torch.empty((n,), device=x.device, dtype=torch.float32)
    if out_idx is None:
        out_idx = torch.empty((n,), device=x.device, dtype=torch.int32)
    L = _native.require()
    rc = L.im_row_argmax(_native.ptr(x), ctypes.c_int(n), ctypes.c_int(c), ctypes.c_int(x.stride(0)),
                         ctypes.c_int(id_offset), _native.ptr(out_val), _native.ptr(out_idx), _native.stream_ptr())
    _native.check(rc, "im_row_argmax")
    _native.count_launch()
    return out_val, out_idx
