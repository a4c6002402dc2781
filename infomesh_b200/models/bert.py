"""BERT-family encoders on the hand-written sm_100a kernels.

Two roles from the reference are served by this one architecture:

* the **embedding encoder** behind ``VectorStore`` (reference infomesh/index/vector_store.py:104-125
  uses ``SentenceTransformer("all-MiniLM-L6-v2")``; the rebuild's flagship stand-in is bge-small-en), and
* the **cross-encoder reranker** that replaces the LLM-prompt reranker
  (reference infomesh/search/reranker.py:86-163) with bge-reranker-base (XLM-R base + classification head).

Weights are random-initialised (seeded) — there is no network access to fetch checkpoints — but the
architecture, shapes and numerics match the named models.  ``forward`` runs the CUDA path
(``ops.gemm`` / ``ops.attention`` / ``ops.nn``); ``forward_ref`` is the plain fp32 PyTorch oracle and
``forward_torch`` the bf16 cuBLAS baseline used by the ``--impl nccl`` bench arm.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class BertConfig:
    name: str = "bge-small-en"
    vocab_size: int = 30522
    hidden: int = 384
    layers: int = 12
    heads: int = 12
    ffn: int = 1536
    max_pos: int = 512
    type_vocab: int = 2
    eps: float = 1e-12
    pos_offset: int = 0          # XLM-R position ids start at padding_idx + 1 = 2
    pooling: str = "cls"         # "cls" | "mean"
    classifier: bool = False     # cross-encoder head (dense + tanh + out_proj -> 1 logit)

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


BGE_SMALL = BertConfig()
MINILM_L6 = BertConfig(name="all-MiniLM-L6-v2", layers=6, pooling="mean")
BGE_RERANKER_BASE = BertConfig(
    name="bge-reranker-base", vocab_size=250002, hidden=768, layers=12, heads=12, ffn=3072, max_pos=514,
    type_vocab=1, eps=1e-5, pos_offset=2, pooling="cls", classifier=True)

CONFIGS = {c.name: c for c in (BGE_SMALL, MINILM_L6, BGE_RERANKER_BASE)}


def _randn(gen, shape, std, device, dtype):
    # generated on CPU for determinism across devices, then moved
    return (torch.randn(shape, generator=gen, dtype=torch.float32) * std).to(device=device, dtype=dtype)


class BertWeights:
    """Parameter container (bf16 matrices, fp32 vectors), optionally a tensor-parallel shard."""

    def __init__(self, cfg: BertConfig, device="cpu", seed: int = 0, tp_rank: int = 0, tp_size: int = 1):
        self.cfg = cfg
        self.tp_rank, self.tp_size = tp_rank, tp_size
        H, F = cfg.hidden, cfg.ffn
        assert cfg.heads % tp_size == 0 and F % tp_size == 0
        g = torch.Generator().manual_seed(seed)
        bf, f32 = torch.bfloat16, torch.float32
        std = 0.02
        self.word = _randn(g, (cfg.vocab_size, H), std, device, bf)
        self.pos = _randn(g, (cfg.max_pos, H), std, device, bf)
        self.type = _randn(g, (cfg.type_vocab, H), std, device, bf)
        self.emb_g = torch.ones(H, device=device, dtype=f32)
        self.emb_b = torch.zeros(H, device=device, dtype=f32)
        self.layers = []
        hs = H // tp_size          # columns of Q/K/V owned by this rank (whole heads)
        fs = F // tp_size
        for _ in range(cfg.layers):
            wq = _randn(g, (H, H), std, "cpu", f32)
            wk = _randn(g, (H, H), std, "cpu", f32)
            wv = _randn(g, (H, H), std, "cpu", f32)
            bqkv = _randn(g, (3, H), std, "cpu", f32)
            wo = _randn(g, (H, H), std, "cpu", f32)
            bo = _randn(g, (H,), std, "cpu", f32)
            w1 = _randn(g, (F, H), std, "cpu", f32)
            b1 = _randn(g, (F,), std, "cpu", f32)
            w2 = _randn(g, (H, F), std, "cpu", f32)
            b2 = _randn(g, (H,), std, "cpu", f32)
            r0, r1 = tp_rank * hs, (tp_rank + 1) * hs
            f0, f1 = tp_rank * fs, (tp_rank + 1) * fs
            lay = dict(
                # column-parallel: rows of the [out, in] matrices
                wqkv=torch.cat([wq[r0:r1], wk[r0:r1], wv[r0:r1]], 0).to(device, bf).contiguous(),
                bqkv=torch.cat([bqkv[0, r0:r1], bqkv[1, r0:r1], bqkv[2, r0:r1]], 0).to(device, f32).contiguous(),
                # row-parallel: columns of the [out, in] matrices; bias added once (rank 0)
                wo=wo[:, r0:r1].to(device, bf).contiguous(),
                bo=(bo if tp_rank == 0 else torch.zeros_like(bo)).to(device, f32),
                ln1_g=torch.ones(H, device=device, dtype=f32), ln1_b=torch.zeros(H, device=device, dtype=f32),
                w1=w1[f0:f1].to(device, bf).contiguous(), b1=b1[f0:f1].to(device, f32).contiguous(),
                w2=w2[:, f0:f1].to(device, bf).contiguous(),
                b2=(b2 if tp_rank == 0 else torch.zeros_like(b2)).to(device, f32),
                ln2_g=torch.ones(H, device=device, dtype=f32), ln2_b=torch.zeros(H, device=device, dtype=f32),
            )
            self.layers.append(lay)
        if cfg.classifier:
            self.cls_w1 = _randn(g, (H, H), std, device, bf)
            self.cls_b1 = torch.zeros(H, device=device, dtype=f32)
            self.cls_w2 = _randn(g, (1, H), std, device, bf)
            self.cls_b2 = torch.zeros(1, device=device, dtype=f32)
            # out_proj padded to one GEMM tile so the head runs on the tensor-core GEMM (column 0 is the logit)
            self.cls_w2p = torch.zeros((128, H), device=device, dtype=bf)
            self.cls_w2p[0] = self.cls_w2[0]
            self.cls_b2p = torch.zeros(128, device=device, dtype=f32)
            self.cls_b2p[0] = self.cls_b2[0]

    def n_params(self) -> int:
        n = self.word.numel() + self.pos.numel() + self.type.numel()
        for lay in self.layers:
            n += sum(v.numel() for v in lay.values())
        return n


class BertModel:
    """Single-GPU (tp_size == 1) encoder / cross-encoder running on the native kernels."""

    def __init__(self, cfg: BertConfig, device="cuda", seed: int = 0, weights: BertWeights | None = None):
        self.cfg = cfg
        self.device = torch.device(device)
        self.w = weights if weights is not None else BertWeights(cfg, device=device, seed=seed)
        # True only when the weights came from a checkpoint (models/loader.py).  Serving code must not let a random-init
        # model influence ranking; benchmarks / tests opt in explicitly.
        self.pretrained = False
        self.source = "random-init"

    # ------------------------------------------------------------------ native path
    def hidden_states(self, ids: torch.Tensor, lengths: torch.Tensor | None = None,
                      type_ids: torch.Tensor | None = None, skip_last: bool = False) -> torch.Tensor:
        """ids: int32 [B, S] -> final hidden states bf16 [B, S, H] (``skip_last``: stop before the last layer)."""
        from infomesh_b200.ops import attention as A
        from infomesh_b200.ops import gemm as G
        from infomesh_b200.ops import nn as N

        cfg, w = self.cfg, self.w
        B, S = ids.shape
        H = cfg.hidden
        x = N.embed_ln(ids.reshape(-1), w.word, w.pos, w.type, w.emb_g, w.emb_b, cfg.eps, S,
                       type_ids=type_ids.reshape(-1) if type_ids is not None else None, pos_offset=cfg.pos_offset)
        for lay in (w.layers[:-1] if skip_last else w.layers):
            qkv = G.linear(x, lay["wqkv"], lay["bqkv"])
            q3 = qkv.view(B, S, 3 * H)
            ctx = A.attention(q3[..., :H], q3[..., H:2 * H], q3[..., 2 * H:], cfg.heads, kv_lens=lengths)
            y = G.linear(ctx.view(B * S, H), lay["wo"], lay["bo"], residual=x)
            x1 = N.layernorm(y, lay["ln1_g"], lay["ln1_b"], cfg.eps)
            h = G.linear(x1, lay["w1"], lay["b1"], act="gelu")
            y2 = G.linear(h, lay["w2"], lay["b2"], residual=x1)
            x = N.layernorm(y2, lay["ln2_g"], lay["ln2_b"], cfg.eps)
        return x.view(B, S, H)

    def hidden_states_packed(self, ids: torch.Tensor, lengths: torch.Tensor, cls_only_last: bool = False,
                             precision: str = "bf16"):
        """Unpadded forward: ids int32 [B, S<=128] + lengths [B] -> (hidden bf16 [B*S, H] of which the first
        ``total`` rows are the packed tokens, cu_seqlens int32 [B+1], total int32 [1]).

        Padding tokens are never embedded, multiplied or normalised: every GEMM / LayerNorm reads the token count
        from a device scalar, and attention walks ``cu_seqlens``.  Shapes stay static, so the whole forward still
        captures into one CUDA graph even though the amount of work follows the batch.  ``cls_only_last`` stops
        before the last layer (see :meth:`_cls_last_layer`)."""
        from infomesh_b200.ops import nn as N

        cfg, w = self.cfg, self.w
        B, S = ids.shape
        H = cfg.hidden
        assert S <= 128 and cfg.head_dim == 64, "packed path: max_seqlen <= 128, head_dim 64"
        pk_ids, pk_pos, cu, total = N.seq_pack(ids.contiguous(), lengths, cfg.pos_offset)
        layers = w.layers[:-1] if cls_only_last else w.layers
        if precision == "mxfp8":
            from infomesh_b200.ops import mx as MX

            xq = MX.alloc_act(B * S, H, ids.device)
            x = N.embed_ln(pk_ids, w.word, w.pos, w.type, w.emb_g, w.emb_b, cfg.eps, S, pos_ids=pk_pos, n_rows_dev=total,
                           mx_out=xq)
            for lay in layers:
                x, xq = self._packed_layer_mx(x, xq, lay, B, S, cu, total)
            return x, cu, total
        x = N.embed_ln(pk_ids, w.word, w.pos, w.type, w.emb_g, w.emb_b, cfg.eps, S, pos_ids=pk_pos, n_rows_dev=total)
        for lay in layers:
            x = self._packed_layer(x, lay, B, S, cu, total, fp8=(precision == "fp8"))
        return x, cu, total

    def _mx_weights(self, lay):
        """MXFP8 copies of a layer's four GEMM weights (e4m3 + ue8m0 per 32, packed once on first use)."""
        from infomesh_b200.ops import mx as MX

        if "wqkv_mx" not in lay:
            for name in ("wqkv", "wo", "w1", "w2"):
                lay[name + "_mx"] = MX.quantize_weight(lay[name])
        return lay

    def _packed_layer_mx(self, x, xq, lay, B, S, cu, total):
        """One encoder layer on block-scaled fp8 GEMMs.  Every GEMM input arrives already quantised by the kernel that
        produced it: LayerNorm (x -> QKV, x1 -> FFN-up), the attention epilogue (context -> out-proj) and the GELU
        epilogue of FFN-up (h -> FFN-down); bf16 survives only on the residual stream and as Q/K/V for attention."""
        from infomesh_b200.ops import attention as A
        from infomesh_b200.ops import mx as MX
        from infomesh_b200.ops import nn as N

        cfg, H = self.cfg, self.cfg.hidden
        self._mx_weights(lay)
        dev = x.device
        qkv = MX.linear_mx(xq, lay["wqkv_mx"], lay["bqkv"], m_dev=total)
        q3 = qkv.view(B, S, 3 * H)
        ctx = MX.alloc_act(B * S, H, dev)
        A.attention_mx(q3[..., :H], q3[..., H:2 * H], q3[..., 2 * H:], cfg.heads, ctx, cu_seqlens=cu)
        x1q = MX.alloc_act(B * S, H, dev)
        x2q = MX.alloc_act(B * S, H, dev)
        if MX.FUSED_LN and H in MX.FUSED_LN_WIDTHS:
            # residual add + LayerNorm + MXFP8 quantisation run in the projection's epilogue (cluster of H / 192 CTAs per row block)
            x1 = MX.linear_mx_ln(ctx, lay["wo_mx"], lay["bo"], x, lay["ln1_g"], lay["ln1_b"], cfg.eps, x1q, m_dev=total)
            h = MX.linear_mx(x1q, lay["w1_mx"], lay["b1"], act="gelu", out_mx=True, m_dev=total)
            x2 = MX.linear_mx_ln(h, lay["w2_mx"], lay["b2"], x1, lay["ln2_g"], lay["ln2_b"], cfg.eps, x2q, m_dev=total)
            return x2, x2q
        y = MX.linear_mx(ctx, lay["wo_mx"], lay["bo"], residual=x, m_dev=total)
        x1 = N.layernorm_mx(y, lay["ln1_g"], lay["ln1_b"], cfg.eps, x1q, n_rows_dev=total)
        h = MX.linear_mx(x1q, lay["w1_mx"], lay["b1"], act="gelu", out_mx=True, m_dev=total)
        y2 = MX.linear_mx(h, lay["w2_mx"], lay["b2"], residual=x1, m_dev=total)
        x2 = N.layernorm_mx(y2, lay["ln2_g"], lay["ln2_b"], cfg.eps, x2q, n_rows_dev=total)
        return x2, x2q

    def _fp8_weights(self, lay):
        """Per-tensor e4m3 copies of a layer's four GEMM weights (made once, on first use of the fp8 path)."""
        from infomesh_b200.ops import gemm as G

        if "wqkv_8" not in lay:
            for name in ("wqkv", "wo", "w1", "w2"):
                lay[name + "_8"], lay[name + "_s"] = G.quantize_weight_fp8(lay[name])
        return lay

    def _lin(self, x, lay, name, bias, fp8, total=None, **kw):
        """One projection: bf16 GEMM, or (fp8) per-token dynamic e4m3 quantisation of the input + kind::f8f6f4 GEMM."""
        from infomesh_b200.ops import gemm as G

        if not fp8:
            return G.linear(x, lay[name], bias, m_dev=total, **kw)
        x8, rs = G.quantize_rows_fp8(x, n_rows_dev=total)
        return G.linear(x8, lay[name + "_8"], bias, alpha=lay[name + "_s"], row_scale=rs, m_dev=total, **kw)

    def _packed_layer(self, x, lay, B, S, cu, total, fp8=False):
        from infomesh_b200.ops import attention as A
        from infomesh_b200.ops import nn as N

        cfg, H = self.cfg, self.cfg.hidden
        if fp8:
            self._fp8_weights(lay)
        qkv = self._lin(x, lay, "wqkv", lay["bqkv"], fp8, total)
        q3 = qkv.view(B, S, 3 * H)
        ctx = A.attention(q3[..., :H], q3[..., H:2 * H], q3[..., 2 * H:], cfg.heads, cu_seqlens=cu)
        y = self._lin(ctx.view(B * S, H), lay, "wo", lay["bo"], fp8, total, residual=x)
        x1 = N.layernorm(y, lay["ln1_g"], lay["ln1_b"], cfg.eps, n_rows_dev=total)
        h = self._lin(x1, lay, "w1", lay["b1"], fp8, total, act="gelu")
        y2 = self._lin(h, lay, "w2", lay["b2"], fp8, total, residual=x1)
        return N.layernorm(y2, lay["ln2_g"], lay["ln2_b"], cfg.eps, n_rows_dev=total)

    def _cls_last_layer(self, x, lay, B, lengths, cu, total):
        """Last encoder layer for a classifier that reads only the <s> token: keys / values for every token, but the
        query, attention output, out-projection, both LayerNorms and the whole FFN only for the B CLS rows (exactly
        the same logits; ~3/4 of the layer's GEMM flops never happen)."""
        from infomesh_b200.ops import attention as A
        from infomesh_b200.ops import gemm as G
        from infomesh_b200.ops import nn as N

        cfg, H = self.cfg, self.cfg.hidden
        kv = G.linear(x, lay["wqkv"][H:], lay["bqkv"][H:], m_dev=total)          # [rows, 2H]: K | V for every token
        x_cls = N.gather_rows(x, cu, B)                                           # [B, H]
        q_cls = G.linear(x_cls, lay["wqkv"][:H], lay["bqkv"][:H])
        kv3 = kv.view(1, kv.shape[0], 2 * H)
        ctx = A.attention_decode(q_cls, kv3[..., :H], kv3[..., H:], cfg.heads, lengths, seq_start=cu)
        y = G.linear(ctx, lay["wo"], lay["bo"], residual=x_cls)
        x1 = N.layernorm(y, lay["ln1_g"], lay["ln1_b"], cfg.eps)
        h = G.linear(x1, lay["w1"], lay["b1"], act="gelu")
        y2 = G.linear(h, lay["w2"], lay["b2"], residual=x1)
        return N.layernorm(y2, lay["ln2_g"], lay["ln2_b"], cfg.eps)               # [B, H] final CLS states

    def score_packed(self, ids, lengths, precision: str = "bf16") -> torch.Tensor:
        """Cross-encoder logits fp32 [B], computed on the unpadded token stream (see :meth:`hidden_states_packed`).

        ``precision="fp8"``: the four projections of layers 0..L-2 run as e4m3 tensor-core GEMMs (per-token dynamic
        activation scales, per-tensor weight scales, fp32 accumulate); attention, norms, the CLS-only last layer and
        the classifier stay bf16.  Opt-in -- the benchmark headline is bf16."""
        from infomesh_b200.ops import gemm as G

        assert self.cfg.classifier
        w = self.w
        x, cu, total = self.hidden_states_packed(ids, lengths, cls_only_last=True, precision=precision)
        cls = self._cls_last_layer(x, w.layers[-1], ids.shape[0], lengths, cu, total)
        hcls = G.linear(cls, w.cls_w1, w.cls_b1, act="tanh")
        return G.linear(hcls, w.cls_w2p, w.cls_b2p, out_dtype=torch.float32)[:, 0].contiguous()

    def embed(self, ids, lengths=None, out_q8=None, out_qscale=None) -> torch.Tensor:
        """Sentence embeddings, L2-normalised bf16 [B, H].  ``out_q8`` / ``out_qscale``: also emit the e4m3 copy + row
        scale the fp8 dense search consumes (same pooling kernel, no extra pass)."""
        from infomesh_b200.ops import nn as N

        h = self.hidden_states(ids, lengths)
        return N.pool_norm(h, lengths, self.cfg.pooling, True, out_q8=out_q8, out_qscale=out_qscale)

    def score(self, ids, lengths=None, type_ids=None) -> torch.Tensor:
        """Cross-encoder relevance logits fp32 [B] on padded ``[B, S]`` batches.  The last encoder layer computes keys /
        values for every token but the query, attention output, out-projection, LayerNorms and FFN only for the <s>
        rows the classifier reads (same logits; see :meth:`_cls_last_layer`)."""
        from infomesh_b200.ops import attention as A
        from infomesh_b200.ops import gemm as G
        from infomesh_b200.ops import nn as N

        assert self.cfg.classifier
        cfg, w, H = self.cfg, self.w, self.cfg.hidden
        B, S = ids.shape
        if cfg.head_dim not in (32, 64) or len(w.layers) < 1:
            return classifier_head(self.hidden_states(ids, lengths, type_ids), w)
        x = self.hidden_states(ids, lengths, type_ids, skip_last=True)               # [B, S, H]
        lay = w.layers[-1]
        kv = G.linear(x.view(B * S, H), lay["wqkv"][H:], lay["bqkv"][H:]).view(B, S, 2 * H)
        x_cls = x[:, 0, :]                                                           # strided rows, no copy
        q_cls = G.linear(x_cls, lay["wqkv"][:H], lay["bqkv"][:H])
        kv_len = lengths if lengths is not None else S
        ctx = A.attention_decode(q_cls, kv[..., :H], kv[..., H:], cfg.heads, kv_len)
        y = G.linear(ctx, lay["wo"], lay["bo"], residual=x_cls)
        x1 = N.layernorm(y, lay["ln1_g"], lay["ln1_b"], cfg.eps)
        h = G.linear(x1, lay["w1"], lay["b1"], act="gelu")
        y2 = G.linear(h, lay["w2"], lay["b2"], residual=x1)
        cls = N.layernorm(y2, lay["ln2_g"], lay["ln2_b"], cfg.eps)
        hcls = G.linear(cls, w.cls_w1, w.cls_b1, act="tanh")
        return G.linear(hcls, w.cls_w2p, w.cls_b2p, out_dtype=torch.float32)[:, 0].contiguous()

    # ------------------------------------------------------------------ references
    def _torch_forward(self, ids, lengths, type_ids, dtype):
        cfg, w = self.cfg, self.w
        B, S = ids.shape
        H, nH, hd = cfg.hidden, cfg.heads, cfg.head_dim
        dev = ids.device
        pos = torch.arange(S, device=dev) + cfg.pos_offset
        ty = type_ids.long() if type_ids is not None else torch.zeros_like(ids, dtype=torch.long)
        x = w.word.to(dtype)[ids.long()] + w.pos.to(dtype)[pos][None] + w.type.to(dtype)[ty]
        x = torch.nn.functional.layer_norm(x.float(), (H,), w.emb_g, w.emb_b, cfg.eps).to(dtype)
        mask = None
        if lengths is not None:
            mask = (torch.arange(S, device=dev)[None] < lengths.view(B, 1))[:, None, None, :]
        for lay in w.layers:
            qkv = x @ lay["wqkv"].to(dtype).t() + lay["bqkv"].to(dtype)
            q, k, v = (t.view(B, S, nH, hd).transpose(1, 2) for t in qkv.split(H, dim=-1))
            ctx = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask)
            ctx = ctx.transpose(1, 2).reshape(B, S, H)
            y = ctx @ lay["wo"].to(dtype).t() + lay["bo"].to(dtype) + x
            x1 = torch.nn.functional.layer_norm(y.float(), (H,), lay["ln1_g"], lay["ln1_b"], cfg.eps).to(dtype)
            h = torch.nn.functional.gelu(x1 @ lay["w1"].to(dtype).t() + lay["b1"].to(dtype))
            y2 = h @ lay["w2"].to(dtype).t() + lay["b2"].to(dtype) + x1
            x = torch.nn.functional.layer_norm(y2.float(), (H,), lay["ln2_g"], lay["ln2_b"], cfg.eps).to(dtype)
        return x

    def _pool(self, h, lengths):
        hf = h.float()
        if self.cfg.pooling == "cls" or lengths is None and self.cfg.pooling == "cls":
            e = hf[:, 0]
        else:
            B, S, _ = hf.shape
            ln = lengths if lengths is not None else torch.full((B,), S, device=h.device)
            m = (torch.arange(S, device=h.device)[None] < ln.view(B, 1)).float()[..., None]
            e = (hf * m).sum(1) / m.sum(1).clamp(min=1)
        return torch.nn.functional.normalize(e, dim=-1)

    def _cls(self, h):
        w = self.w
        z = torch.tanh(h[:, 0].float() @ w.cls_w1.float().t() + w.cls_b1)
        return (z @ w.cls_w2.float().t()).squeeze(1) + w.cls_b2

    def hidden_states_ref(self, ids, lengths=None, type_ids=None):
        """fp32 PyTorch oracle."""
        return self._torch_forward(ids, lengths, type_ids, torch.float32)

    def embed_ref(self, ids, lengths=None):
        return self._pool(self.hidden_states_ref(ids, lengths), lengths)

    def score_ref(self, ids, lengths=None, type_ids=None):
        return self._cls(self.hidden_states_ref(ids, lengths, type_ids))

    def embed_torch(self, ids, lengths=None):
        """bf16 cuBLAS / SDPA baseline (the "reference's own PyTorch build")."""
        return self._pool(self._torch_forward(ids, lengths, None, torch.bfloat16), lengths).to(torch.bfloat16)

    def score_torch(self, ids, lengths=None, type_ids=None):
        return self._cls(self._torch_forward(ids, lengths, type_ids, torch.bfloat16))

    def flops_per_token(self, seq_len: int) -> float:
        cfg = self.cfg
        H, F = cfg.hidden, cfg.ffn
        per_layer = 2 * (3 * H * H + H * H + 2 * H * F) + 4 * seq_len * H
        return float(cfg.layers * per_layer)


def classifier_head(h: torch.Tensor, w: "BertWeights") -> torch.Tensor:
    """RoBERTa-style head on the <s> token: dense + tanh -> out_proj, as two tensor-core GEMMs over the strided CLS rows
    (the per-sequence ``cls_head`` kernel re-read the 768x768 dense matrix once per sequence: 290 us at B = 1280)."""
    from infomesh_b200.ops import gemm as G

    x = G.linear(h[:, 0, :], w.cls_w1, w.cls_b1, act="tanh")
    return G.linear(x, w.cls_w2p, w.cls_b2p, out_dtype=torch.float32)[:, 0].contiguous()


def param_count(cfg: BertConfig) -> int:
    H, F = cfg.hidden, cfg.ffn
    per = 4 * H * H + 2 * H * F + 4 * H + F + H + 4 * H
    return cfg.vocab_size * H + cfg.max_pos * H + cfg.type_vocab * H + cfg.layers * per


__all__ = ["BertConfig", "BertModel", "BertWeights", "BGE_SMALL", "MINILM_L6", "BGE_RERANKER_BASE", "CONFIGS",
           "param_count"]
_ = math
