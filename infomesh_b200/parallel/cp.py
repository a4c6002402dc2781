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
