"""Context parallelism for long inputs (SURVEY §2.3 P-CP / §5.7): the sequence is sharded over the ranks and attention
runs either Ulysses-style (all-to-all: sequence-sharded -> head-sharded, local full-sequence attention, all-to-all
back) or as an all-pairs K/V exchange with log-sum-exp merging (the NVSwitch form of ring attention: every peer is one
hop away, so all K/V blocks are fetched at once instead of being passed around a ring).

The reference has nothing comparable -- it truncates (2000 chars for embeddings, 8000 for summaries:
infomesh/index/vector_store.py:142-156, infomesh/summarizer/engine.py:382,396).  This is the optional path for
summarising / embedding a full ~100 KB document untruncated; the default pipeline keeps the truncation semantics.

Data movement uses ``torch.distributed`` collectives (NCCL on GPUs, gloo on CPU for the tests); the attention itself is
the fused kernel on CUDA tensors and the fp32 oracle otherwise."""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def _attend(q, k, v, n_heads, kv_lens=None, scale=None):
    """[B, Sq, h*d] x [B, Sk, h*d] -> [B, Sq, h*d]; fused kernel on CUDA bf16, oracle elsewhere."""
    from infomesh_b200.ops import attention as A

    if q.is_cuda and q.dtype == torch.bfloat16:
        return A.attention(q.contiguous(), k.contiguous(), v.contiguous(), n_heads, kv_lens=kv_lens, scale=scale)
    return A.attention_ref(q, k, v, n_heads, kv_lens=kv_lens, scale=scale).to(q.dtype)


def ulysses_attention(q, k, v, n_heads: int, group=None, scale: float | None = None):
    """Sequence-sharded self-attention via head scatter.

    ``q, k, v``: this rank's sequence shard ``[B, S/W, n_heads*d]`` (rank r holds positions ``[r*S/W, (r+1)*S/W)``).
    Requires ``n_heads % W == 0``.  Returns the attention output for the same shard."""
    W = dist.get_world_size(group) if dist.is_initialized() else 1
    if W == 1:
        return _attend(q, k, v, n_heads, scale=scale)
    B, s_loc, HH = q.shape
    assert n_heads % W == 0, "Ulysses needs the head count divisible by the group size"
    d = HH // n_heads
    h_loc = n_heads // W

    def seq_to_heads(x):
        # [B, s_loc, W, h_loc*d] -> send head-group w to rank w; receive every rank's sequence shard of my head group
        xs = x.reshape(B, s_loc, W, h_loc * d).permute(2, 0, 1, 3).contiguous()          # [W, B, s_loc, h_loc*d]
        out = torch.empty_like(xs)
        dist.all_to_all_single(out, xs, group=group)
        return out.permute(1, 0, 2, 3).reshape(B, W * s_loc, h_loc * d)                  # full sequence, my heads

    def heads_to_seq(x):
        xs = x.reshape(B, W, s_loc, h_loc * d).permute(1, 0, 2, 3).contiguous()          # [W(seq shard), B, s_loc, hd]
        out = torch.empty_like(xs)
        dist.all_to_all_single(out, xs, group=group)                                     # [W(head group), B, s_loc, hd]
        return out.permute(1, 2, 0, 3).reshape(B, s_loc, HH)

    o = _attend(seq_to_heads(q), seq_to_heads(k), seq_to_heads(v), h_loc, scale=scale)
    return heads_to_seq(o)


def _attend_lse(q, k, v, n_heads, scale):
    """fp32 attention that also returns the per-(row, head) log-sum-exp, for merging partial results."""
    B, Sq, HH = q.shape
    d = HH // n_heads
    qf = q.float().view(B, Sq, n_heads, d).transpose(1, 2)
    kf = k.float().view(B, -1, n_heads, d).transpose(1, 2)
    vf = v.float().view(B, -1, n_heads, d).transpose(1, 2)
    s = (qf @ kf.transpose(-1, -2)) * scale
    lse = torch.logsumexp(s, dim=-1)                                                     # [B, h, Sq]
    o = torch.softmax(s, dim=-1) @ vf
    return o, lse


def allpairs_attention(q, k, v, n_heads: int, group=None, scale: float | None = None):
    """Sequence-sharded self-attention where every rank gathers all K/V shards (one all-gather: on NVSwitch every peer
    is one hop, so there is no ring to walk) and merges the per-shard partial outputs with their log-sum-exp weights --
    numerically the blockwise / ring-attention recurrence.  Works for any head count; K/V memory is ``S`` per rank,
    scores are never larger than ``[S/W, S/W]`` per block."""
    W = dist.get_world_size(group) if dist.is_initialized() else 1
    B, s_loc, HH = q.shape
    d = HH // n_heads
    scale = (1.0 / math.sqrt(d)) if scale is None else scale
    if W == 1:
        ks, vs = [k], [v]
    else:
        ks = [torch.empty_like(k) for _ in range(W)]
        vs = [torch.empty_like(v) for _ in range(W)]
        dist.all_gather(ks, k.contiguous(), group=group)
        dist.all_gather(vs, v.contiguous(), group=group)
    acc, lse_run = None, None
    for kb, vb in zip(ks, vs):
        o, lse = _attend_lse(q, kb, vb, n_heads, scale)
        if acc is None:
            acc, lse_run = o, lse
        else:
            new = torch.logaddexp(lse_run, lse)
            acc = acc * torch.exp(lse_run - new)[..., None] + o * torch.exp(lse - new)[..., None]
            lse_run = new
    return acc.transpose(1, 2).reshape(B, s_loc, HH).to(q.dtype)


# =====================================================================================================================
# In-kernel context parallelism (the product path; the functions above are the torch.distributed baseline / CPU oracle)
# =====================================================================================================================
class CpQkvBuffers:
    """Double-buffered ``[B * S_local, 3 * inner]`` QKV activations in the symmetric heap.

    The QKV projection of layer ``l`` writes buffer ``l % 2``; after ONE heap barrier every rank's attention kernel pulls
    the K / V tiles of all ranks straight out of those buffers (``attn_fwd_kernel<.., CP>``: TMA loads through tensor
    maps over the peers' mapped memory, all-pairs order starting with the own shard).  Alternating the buffer makes that
    single barrier sufficient: a rank reaches the barrier of layer ``l + 1`` only after its layer-``l`` attention has
    finished reading, so nobody overwrites buffer ``l % 2`` (layer ``l + 2``) while a peer still reads it."""

    def __init__(self, heap, batch: int, s_local: int, inner: int):
        import ctypes

        self.heap, self.B, self.s_local, self.inner = heap, batch, s_local, inner
        c = heap.ctx
        self.world, self.rank = c.world, c.rank
        self.bufs, self._k_ptrs, self._v_ptrs = [], [], []
        for _ in range(2):
            view, off = heap.alloc((batch * s_local, 3 * inner), torch.bfloat16)
            self.bufs.append(view)
            self._k_ptrs.append((ctypes.c_void_p * c.world)(*[b + off + inner * 2 for b in heap.bases]))
            self._v_ptrs.append((ctypes.c_void_p * c.world)(*[b + off + 2 * inner * 2 for b in heap.bases]))


def cp_attention(bufs: CpQkvBuffers, parity: int, n_heads: int, *, kv_lens=None, scale: float | None = None, rel_bias=None, out=None):
    """Attention of this rank's queries (``bufs.bufs[parity][:, :inner]``) over the keys / values of every rank.

    ``kv_lens`` int32 ``[B]``: GLOBAL valid key counts; ``rel_bias`` fp32 ``[heads, 2 * S_total - 1]`` natural-log table of
    the whole sequence (T5).  The caller must have issued ``heap.barrier()`` after the projection that filled the buffer."""
    import ctypes

    from infomesh_b200 import _native
    from infomesh_b200.ops.attention import LOG2E

    qkv = bufs.bufs[parity]
    inner, B, s_loc = bufs.inner, bufs.B, bufs.s_local
    hd = inner // n_heads
    scale = (1.0 / math.sqrt(hd)) if scale is None else scale
    if out is None:
        out = torch.empty((B * s_loc, inner), device=qkv.device, dtype=torch.bfloat16)
    bias_dev = (rel_bias.float() * LOG2E).contiguous() if rel_bias is not None else None
    if bias_dev is not None:
        assert bias_dev.shape == (n_heads, 2 * s_loc * bufs.world - 1)
    L = _native.require()
    rc = L.im_attn_fwd_cp(_native.ptr(qkv), bufs._k_ptrs[parity], bufs._v_ptrs[parity], _native.ptr(out), ctypes.c_int(B),
                          ctypes.c_int(n_heads), ctypes.c_int(hd), ctypes.c_int(s_loc), ctypes.c_int(s_loc), ctypes.c_int(bufs.world),
                          ctypes.c_int(bufs.rank), ctypes.c_int(3 * inner), ctypes.c_int(3 * inner), ctypes.c_int(3 * inner),
                          ctypes.c_int(out.stride(0)), _native.ptr(kv_lens), ctypes.c_float(scale), _native.ptr(bias_dev),
                          _native.stream_ptr())
    _native.check(rc, "im_attn_fwd_cp")
    _native.count_launch()
    bufs._keep_bias = bias_dev
    return out


class CPT5Encoder:
    """T5 encoder over a sequence sharded across the ranks (rank ``r`` holds positions ``[r * S/W, (r + 1) * S/W)`` of every
    sequence): embeddings, norms, projections and the FFN are row-local with replicated weights; only attention crosses
    ranks, and it does so INSIDE the attention kernel (:func:`cp_attention`).  Lets the summariser read a full ~100 KB
    page (the reference truncates at 8000 characters, infomesh/summarizer/engine.py:382,396)."""

    def __init__(self, model, batch: int, s_total: int, heap=None):
        from infomesh_b200.parallel import dist as D
        from infomesh_b200.parallel import symm

        self.model, self.ctx = model, D.ctx()
        W = self.ctx.world
        assert s_total % (128 * W) == 0, "sequence length must split into 128-key chunks per rank"
        self.B, self.S, self.s_local = batch, s_total, s_total // W
        cfg = model.cfg
        need = 2 * batch * self.s_local * 3 * cfg.inner * 2 + (1 << 20)
        self.heap = heap or symm.SymmetricHeap(need + (4 << 20), self.ctx)
        self.bufs = CpQkvBuffers(self.heap, batch, self.s_local, cfg.inner)
        self.heap.barrier()

    def encode_local(self, ids_local: torch.Tensor, lengths: torch.Tensor | None = None) -> torch.Tensor:
        """``ids_local`` int32 ``[B, S/W]`` (this rank's slice) -> this rank's encoder states bf16 ``[B, S/W, d_model]``."""
        from infomesh_b200.ops import gemm as G
        from infomesh_b200.ops import nn as N

        m, cfg, w = self.model, self.model.cfg, self.model.w
        B, s_loc, inner = self.B, self.s_local, cfg.inner
        x = w.emb[ids_local.reshape(-1).long()]
        bias = m._enc_bias_table(self.S)
        for li, lay in enumerate(w.enc):
            par = li & 1
            n1 = N.layernorm(x, lay["ln1"], None, cfg.eps, rms_only=True)
            G.linear(n1, lay["wqkv"], out=self.bufs.bufs[par])
            self.heap.barrier()                                   # every rank's K / V of this layer are in place
            ctx = cp_attention(self.bufs, par, cfg.heads, kv_lens=lengths, scale=1.0, rel_bias=bias)
            x = G.linear(ctx, lay["wo"], residual=x)
            n2 = N.layernorm(x, lay["ln2"], None, cfg.eps, rms_only=True)
            h = G.linear(n2, lay["wi"], act="relu")
            x = G.linear(h, lay["wo2"], residual=x)
        return N.layernorm(x, w.enc_final, None, cfg.eps, rms_only=True).view(B, s_loc, cfg.d_model)

    def close(self):
        self.heap.close()
