"""Tensor-parallel + sequence-parallel BERT-class encoder with the collectives fused into the compute kernels
(SURVEY P-TP / P-SP, kernels K6 and K8).

Layout (Megatron-SP): between blocks the activations are sharded by *rows* (each rank owns ``M / tp`` token rows:
LayerNorm / residual run on the shard); inside a block the weights are sharded by columns (QKV, FFN-up) or rows
(out-proj, FFN-down).  The two boundaries are an all-gather and a reduce-scatter, and neither is a separate call:

    sum_ln (LN + residual)  ── writes its normalised rows straight into every peer's full-sequence buffer and bumps
                               the per-128-row arrival counters there                        (all-gather, producer)
    column-parallel GEMM    ── TMA producer waits per row block, own rows first (m_rotate)   (all-gather, consumer)
    row-parallel GEMM       ── epilogue pushes each 128-row block of the partial product to the rank that owns those
                               rows, then bumps that rank's arrival counter                   (reduce-scatter, producer)
    sum_ln                  ── waits on the counters, sums the tp partials + residual, LN     (reduce-scatter, consumer)

so a layer is the same 7 launches as the single-GPU model and the NVLink traffic overlaps the tensor-core tiles of the
same kernel.  ``comm="nccl"`` runs the identical math with ``all_gather_into_tensor`` / ``reduce_scatter_tensor``
between plain kernels: the baseline for A/B.  The reference has no model parallelism at all (it calls
``SentenceTransformer.encode`` on one device, infomesh/index/vector_store.py:104-125).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from infomesh_b200.models.bert import BertConfig, BertWeights, classifier_head
from infomesh_b200.ops import attention as A
from infomesh_b200.ops import gemm as G
from infomesh_b200.ops import nn as N
from infomesh_b200.parallel import dist as D
from infomesh_b200.parallel.symm import AllGatherChannel, SymmetricHeap


def _bn_for(n_cols: int, m_rows: int) -> int:
    """Mirror of the GEMM launcher's tile choice so the consumer knows how many epilogue arrivals to expect."""
    tiles256 = ((m_rows + 127) // 128) * ((n_cols + 255) // 256)
    return 256 if (n_cols % 256 == 0 and tiles256 >= 148) else 128


class ReduceScatterChannel:
    """Receive slots ``[tp_src][rows_per_rank][n_cols]`` + arrival counters ``[tp_src][rows_per_rank / 128]`` on every
    rank; producers are row-parallel GEMM epilogues, the consumer is ``sum_ln``."""

    def __init__(self, heap: SymmetricHeap, rows_per_rank: int, n_cols: int):
        c = heap.ctx
        assert rows_per_rank % 128 == 0, "sequence-parallel shards must be multiples of 128 rows"
        self.world, self.rank, self.rows_per_rank, self.n_cols = c.world, c.rank, rows_per_rank, n_cols
        self.blocks_per_src = rows_per_rank // 128
        self.recv, ro = heap.alloc((c.world, rows_per_rank, n_cols), torch.bfloat16)
        self.flags, fo = heap.alloc((c.world, self.blocks_per_src), torch.int32)
        self._c_tab, self._f_tab = heap.peer_table(ro), heap.peer_table(fo)
        self.peer_c_ptr, self.peer_flags_ptr = self._c_tab.data_ptr(), self._f_tab.data_ptr()
        import ctypes

        self.peer_c_host = (ctypes.c_void_p * c.world)(*[b + ro for b in heap.bases])   # host copy: TMA store maps
        self.local_flags_ptr = self.flags.data_ptr()
        self.state = torch.zeros(2, dtype=torch.int32, device=c.device)
        self.step_ptr = self.state.data_ptr()
        self.bn = _bn_for(n_cols, rows_per_rank * c.world)
        self.arrivals_per_block = 8 * ((n_cols + self.bn - 1) // self.bn)     # 8 epilogue warps x N tiles


class _AgPush:
    def __init__(self, peer_buf_ptr, peer_flags_ptr, row_offset, world, rank):
        self.peer_buf_ptr, self.peer_flags_ptr, self.row_offset, self.world, self.rank = peer_buf_ptr, peer_flags_ptr, row_offset, world, rank


class AllGatherInput:
    """Full-sequence buffer ``[M, H]`` that peers fill row by row, with per-row-block arrival counters."""

    def __init__(self, heap: SymmetricHeap, m_rows: int, n_cols: int, rows_per_rank: int):
        c = heap.ctx
        self.world, self.rank = c.world, c.rank
        self.buf, bo = heap.alloc((m_rows, n_cols), torch.bfloat16)
        nblk = (m_rows + 127) // 128
        self.flags, fo = heap.alloc((nblk,), torch.int32)
        self._b_tab, self._f_tab = heap.peer_table(bo), heap.peer_table(fo)
        self.flags_ptr = self.flags.data_ptr()
        self.state = torch.zeros(2, dtype=torch.int32, device=c.device)
        self.state_ptr = self.state.data_ptr()
        self.m_rotate = (c.rank * rows_per_rank) // 128
        self.push = _AgPush(self._b_tab.data_ptr(), self._f_tab.data_ptr(), c.rank * rows_per_rank, c.world, c.rank)


class TPBertModel:
    """``tp`` = world size of the default group.  Inputs (ids, lengths) are replicated; rank r owns sequences
    ``[r * B/tp, (r+1) * B/tp)`` of the residual stream."""

    def __init__(self, cfg: BertConfig, batch: int, seq_len: int, *, heap: SymmetricHeap | None = None, seed: int = 0, comm: str = "fused"):
        self.ctx = D.ctx()
        c = self.ctx
        self.cfg, self.comm, self.B, self.S = cfg, comm, batch, seq_len
        self.tp = c.world
        assert batch % self.tp == 0 and cfg.heads % self.tp == 0
        self.rows = batch // self.tp * seq_len
        self.M = batch * seq_len
        self.w = BertWeights(cfg, device=c.device, seed=seed, tp_rank=c.rank, tp_size=self.tp)
        H = cfg.hidden
        if comm == "fused":
            need = 2 * self.M * H * 2 + 2 * self.tp * self.rows * H * 2 + (1 << 20)
            self.heap = heap or SymmetricHeap(need + (8 << 20), c)
            self.ag1 = AllGatherInput(self.heap, self.M, H, self.rows)
            self.ag2 = AllGatherInput(self.heap, self.M, H, self.rows)
            self.rs1 = ReduceScatterChannel(self.heap, self.rows, H)
            self.rs2 = ReduceScatterChannel(self.heap, self.rows, H)
            n_log = (batch // self.tp + 3) // 4 * 4
            self.ag_logits = AllGatherChannel(self.heap, (n_log,), torch.float32)
            self._log_stage = torch.zeros(n_log, dtype=torch.float32, device=c.device)
            self.heap.barrier()
        else:
            self.heap = None

    # ------------------------------------------------------------------ forward
    def hidden_states_local(self, ids: torch.Tensor, lengths: torch.Tensor | None = None) -> torch.Tensor:
        """-> this rank's rows of the final hidden states, bf16 ``[B/tp, S, H]``."""
        cfg, w, c = self.cfg, self.w, self.ctx
        B, S, H, tp = self.B, self.S, cfg.hidden, self.tp
        assert tuple(ids.shape) == (B, S)
        # ids are replicated, so every rank embeds the whole batch itself: ~0.2 ms of redundant gather+LN at B=1280
        # instead of a [M, H] all-gather in front of the first QKV GEMM
        x_full = N.embed_ln(ids.reshape(-1).contiguous(), w.word, w.pos, w.type, w.emb_g, w.emb_b, cfg.eps, S,
                            pos_offset=cfg.pos_offset)
        r0 = c.rank * self.rows
        x_loc = x_full[r0:r0 + self.rows]
        hs, heads = H // tp, cfg.heads // tp
        fused = self.comm == "fused"
        n_layers = len(w.layers)
        for li, lay in enumerate(w.layers):
            # ---- attention block: column-parallel QKV (all-gather consumer) -> local heads -> row-parallel out-proj
            qkv = G.linear(x_full, lay["wqkv"], lay["bqkv"], ag=self.ag1 if (fused and li > 0) else None).view(B, S, 3 * hs)
            ctx_ = A.attention(qkv[..., :hs], qkv[..., hs:2 * hs], qkv[..., 2 * hs:], heads, kv_lens=lengths)
            if fused:
                G.linear(ctx_.view(self.M, hs), lay["wo"], lay["bo"], rs=self.rs1, bn=self.rs1.bn)
                # the normalised rows land straight in every rank's full-sequence buffer (own copy included): the local
                # shard used as the next residual is a view of that buffer, so there is no separate local write
                N.layernorm(self.rs1.recv.view(tp, self.rows, H), lay["ln1_g"], lay["ln1_b"], cfg.eps, residual=x_loc, partials=tp,
                            partial_stride=self.rows * H, rs=self.rs1, ag_push=self.ag2.push, want_norm=False)
                x_full2 = self.ag2.buf
                x1 = x_full2[r0:r0 + self.rows]
            else:
                part = G.linear(ctx_.view(self.M, hs), lay["wo"], lay["bo"])
                x1 = N.layernorm(self._nccl_rs(part), lay["ln1_g"], lay["ln1_b"], cfg.eps, residual=x_loc)
                x_full2 = self._nccl_ag(x1)
            # ---- FFN block
            h = G.linear(x_full2, lay["w1"], lay["b1"], act="gelu", ag=self.ag2 if fused else None)
            last = li == n_layers - 1
            if fused:
                G.linear(h, lay["w2"], lay["b2"], rs=self.rs2, bn=self.rs2.bn)
                x_loc = N.layernorm(self.rs2.recv.view(tp, self.rows, H), lay["ln2_g"], lay["ln2_b"], cfg.eps, residual=x1, partials=tp,
                                    partial_stride=self.rows * H, rs=self.rs2, ag_push=None if last else self.ag1.push,
                                    want_norm=last)
                x_full = self.ag1.buf
                if not last:
                    x_loc = x_full[r0:r0 + self.rows]
            else:
                part = G.linear(h, lay["w2"], lay["b2"])
                x_loc = N.layernorm(self._nccl_rs(part), lay["ln2_g"], lay["ln2_b"], cfg.eps, residual=x1)
                x_full = x_loc if last else self._nccl_ag(x_loc)
        return x_loc.view(B // tp, S, H)

    def score(self, ids: torch.Tensor, lengths: torch.Tensor | None = None) -> torch.Tensor:
        """Cross-encoder logits fp32 ``[B]`` on every rank."""
        assert self.cfg.classifier
        w = self.w
        h = self.hidden_states_local(ids, lengths)
        local = classifier_head(h, w)
        n = local.numel()
        if self.comm == "fused":
            self._log_stage[:n].copy_(local)
            return self.ag_logits(self._log_stage)[:, :n].reshape(-1)
        return D.all_gather_cat(local.contiguous()).reshape(-1)

    def hidden_states(self, ids: torch.Tensor, lengths: torch.Tensor | None = None) -> torch.Tensor:
        """Full ``[B, S, H]`` on every rank (testing / pooling)."""
        loc = self.hidden_states_local(ids, lengths).contiguous()
        return D.all_gather_cat(loc).reshape(self.B, self.S, self.cfg.hidden)

    # ------------------------------------------------------------------ NCCL baseline collectives
    def _nccl_ag(self, x_loc: torch.Tensor) -> torch.Tensor:
        out = torch.empty((self.M, x_loc.shape[1]), device=x_loc.device, dtype=x_loc.dtype)
        dist.all_gather_into_tensor(out, x_loc.contiguous())
        return out

    def _nccl_rs(self, part: torch.Tensor) -> torch.Tensor:
        out = torch.empty((self.rows, part.shape[1]), device=part.device, dtype=part.dtype)
        dist.reduce_scatter_tensor(out, part.contiguous())
        return out
