"""The hybrid search pipeline on GPU — one fused device pass per query batch (SURVEY.md §3.2 mapping).

    tokens ──► encoder (bge-small) ──► q_emb ─┬─► sim_topk over the local dense shard ─┐
    terms  ───────────────────────────────────┴─► BM25 AND/score/top-k over local CSR ─┤  per-shard lists
                                                        NVLink exchange + topk_merge ◄──┘
                               RRF fuse (k = 60) ──► build <s> q </s></s> passage </s> pairs
                               ──► cross-encoder (bge-reranker-base) ──► rerank_select ──► top-10 ids / scores

It replaces ``search_hybrid`` + ``rerank_with_llm`` of the reference (infomesh/search/query.py:244-319,
infomesh/search/reranker.py:86-163).  Every rank holds a document-partitioned shard (dense vectors, postings,
passage tokens); all ranks see the same query batch; the reranker is data-parallel over queries.
``backend="fused"`` runs the hand-written kernels, ``backend="torch"`` is the PyTorch (cuBLAS/SDPA + NCCL)
baseline build of the same pipeline used for A/B.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from infomesh_b200.models.bert import BGE_RERANKER_BASE, BGE_SMALL, BertModel
from infomesh_b200.ops import fuse as F
from infomesh_b200.ops import search as S
from infomesh_b200.parallel import dist as D


@dataclass
class HybridConfig:
    nq: int = 64              # queries per batch
    enc_seq: int = 32         # encoder sequence length (query tokens incl. specials, padded)
    max_q_tokens: int = 32    # reranker-side query tokens
    max_terms: int = 8        # BM25 query terms
    k_fetch: int = 20         # candidates per source (reference fetches limit*2, query.py:124)
    n_rerank: int = 20        # reference reranks <= 20 candidates (reranker.py:20)
    k_out: int = 10
    varlen: bool = True       # cross-encoder runs on the unpadded token stream (padding never reaches a kernel)
    precision: str = "bf16"   # cross-encoder GEMMs: "bf16" | "mxfp8" (block-scaled e4m3, quantisers fused into the
                              # producers) | "fp8" (round-1 per-row e4m3 with a standalone quantiser; kept for A/B)
    rerank_chunks: int = 1    # split the rank's pairs into this many sub-batches so activations stay L2-resident
    pair_seq: int = 128       # <s> q </s></s> passage </s>
    rerank: bool = True
    dense: bool = True        # False: BM25-only retrieval (serving with an encoder that has no checkpoint weights)
    dense_dtype: str = "bf16"  # "fp8": e4m3 shard + row scales, 32-wide over-fetch re-scored against the bf16 rows (K3)
    shard_encoder: bool = True  # multi-GPU + p2p exchange: each rank encodes nq / world queries and the embeddings are all-gathered
                                # through the symmetric heap (one push kernel) instead of every rank encoding the whole batch
    retrieval_sms: int = 0     # pipelined serving: SMs left to the HBM-bound scan while the cross-encoder GEMMs of the previous
                               # batch run on the others (0 = no partition: the two streams time-share the whole GPU)
    rank_signals: bool = False  # BM25-only mode: order candidates with the six-signal rank fuse (K12) instead of raw BM25
    backend: str = "fused"    # "fused" | "torch"
    use_graph: bool = True
    exchange: str = "p2p"     # multi-GPU list exchange: "p2p" (fused peer-memory kernels) | "nccl" (baseline collectives)
    strict_graph: bool = False  # CUDA-graph capture failure is an error (benchmarks) instead of a warning + eager fallback
    degraded_ok: bool = False  # p2p exchange: drop a silent shard after ``wait_limit`` polls instead of trapping
    wait_limit: int = 0        # 0 = library default (~1 s)


class HybridEngine:
    def __init__(self, shard, cfg: HybridConfig, encoder: BertModel | None = None, reranker: BertModel | None = None,
                 passage_tables=None, docs_per_shard: int | None = None, seed: int = 0, passage_len: int | None = None):
        self.cfg = cfg
        self.ctx = D.ctx()
        self.shard = shard
        dev = shard.device
        self.device = dev
        self.encoder = encoder or BertModel(BGE_SMALL, device=dev, seed=seed + 1)
        self.reranker = (reranker or BertModel(BGE_RERANKER_BASE, device=dev, seed=seed + 2)) if cfg.rerank else None
        w = self.ctx.world
        assert cfg.nq % w == 0, "query batch must divide evenly over ranks for the data-parallel reranker"
        self.nq_local = cfg.nq // w
        # passage-token tables: one (tok, len) pointer pair per shard.  Default: this rank's own shard holds the
        # tokens of every document it may need (replicated store); a symmetric heap supplies peer pointers.
        # A (tok_ptrs, len_ptrs) pair of int64 device tensors is a ready-made PEER table: entry p points at rank p's
        # passage shard in the symmetric heap, so pair assembly pulls the winning passages straight out of the owning
        # GPU's HBM over NVLink (no replicated passage store).
        if passage_tables is None:
            passage_tables = ([shard.passage_tok], [shard.passage_len])
            docs_per_shard = shard.passage_tok.shape[0]
        if isinstance(passage_tables[0], torch.Tensor):
            self.tok_ptrs, self.len_ptrs = passage_tables[0], passage_tables[1]
            self.passage_len = int(passage_len if passage_len is not None else shard.passage_tok.shape[1])
        else:
            self.tok_ptrs = F.ptr_table(passage_tables[0], dev)
            self.len_ptrs = F.ptr_table(passage_tables[1], dev)
            self.passage_len = passage_tables[0][0].shape[1]
        self._keep = passage_tables
        self.docs_per_shard = int(docs_per_shard)
        # static I/O buffers (CUDA-graph friendly)
        i32 = dict(device=dev, dtype=torch.int32)
        self.in_enc_ids = torch.zeros((cfg.nq, cfg.enc_seq), **i32)
        self.in_enc_len = torch.ones((cfg.nq,), **i32)
        self.in_q_tok = torch.ones((cfg.nq, cfg.max_q_tokens), **i32)
        self.in_q_len = torch.ones((cfg.nq,), **i32)
        self.in_terms = torch.full((cfg.nq, cfg.max_terms), -1, **i32)
        self.out_scores = torch.zeros((cfg.nq, cfg.k_out), device=dev, dtype=torch.float32)
        self.out_ids = torch.full((cfg.nq, cfg.k_out), -1, device=dev, dtype=torch.int64)
        self._f8 = cfg.dense and cfg.dense_dtype == "fp8" and cfg.backend == "fused"
        if self._f8:
            if getattr(shard, "vectors_f8", None) is None:
                from infomesh_b200.ops import nn as N

                shard.vectors_f8, shard.vec_scale = N.quantize_rows_e4m3(shard.vectors)
            self.q8 = torch.zeros((cfg.nq, shard.vectors.shape[1]), device=dev, dtype=torch.uint8)
            self.q8_scale = torch.ones((cfg.nq,), device=dev, dtype=torch.float32)
        self.ch_qemb = None
        self._graph = None
        self._graph_failed = False
        self.heap = None
        self._inject_drop = False
        if self.ctx.is_dist and cfg.backend == "fused" and cfg.exchange == "p2p":
            from infomesh_b200.parallel import symm

            self.heap = symm.SymmetricHeap(8 << 20, self.ctx)
            kw = dict(degraded_ok=cfg.degraded_ok, wait_limit=cfg.wait_limit)
            self.ch_dense = symm.TopkChannel(self.heap, cfg.nq, cfg.k_fetch, **kw)
            self.ch_bm25 = symm.TopkChannel(self.heap, cfg.nq, cfg.k_fetch, **kw)
            import os

            self._inject_drop = os.environ.get("INFOMESH_B200_INJECT_DROP_RANK", "") == str(self.ctx.rank)   # fault-injection hook
            n_log = self.nq_local * cfg.n_rerank
            self._log_pad = (n_log + 3) // 4 * 4           # 16-byte blocks
            self.ch_logits = symm.AllGatherChannel(self.heap, (self._log_pad,), torch.float32)
            dim = shard.vectors.shape[1]
            if cfg.shard_encoder and (self.nq_local * dim * 2) % 16 == 0:
                self.ch_qemb = symm.AllGatherChannel(self.heap, (self.nq_local, dim), torch.bfloat16)
                self._q_scratch = torch.empty((cfg.nq, dim), device=dev, dtype=torch.bfloat16)
            self._log_stage = torch.zeros((self._log_pad,), device=dev, dtype=torch.float32)
            self.heap.barrier()

    # ------------------------------------------------------------------ stages
    def _encode(self):
        if self.cfg.backend == "torch":
            return self.encoder.embed_torch(self.in_enc_ids, self.in_enc_len)
        if self.ch_qemb is not None:
            # de-replicated query encoder: this rank's slice of the batch, then a push all-gather of the [nq/world, dim]
            # embeddings (a few KB per rank) -- every rank ends up with the same [nq, dim] matrix the replicated form computes
            lo = self.ctx.rank * self.nq_local
            mine = self.encoder.embed(self.in_enc_ids[lo:lo + self.nq_local], self.in_enc_len[lo:lo + self.nq_local])
            q_emb = self.ch_qemb(mine.contiguous()).view(self.cfg.nq, -1)
            if self._f8:     # e4m3 copy + scales of the gathered matrix (the pooling kernel, sequence length 1)
                from infomesh_b200.ops import nn as N

                N.pool_norm(q_emb.view(self.cfg.nq, 1, -1), None, "cls", False, out=self._q_scratch, out_q8=self.q8, out_qscale=self.q8_scale)
            return q_emb
        if self._f8:
            return self.encoder.embed(self.in_enc_ids, self.in_enc_len, out_q8=self.q8, out_qscale=self.q8_scale)
        return self.encoder.embed(self.in_enc_ids, self.in_enc_len)

    def _dense_local(self, q_emb):
        cfg, sh = self.cfg, self.shard
        if cfg.backend == "torch":
            sc = q_emb @ sh.vectors.t()
            v, i = torch.topk(sc.float(), min(cfg.k_fetch, sh.vectors.shape[0]), dim=1)
            return v, i.long() + sh.cfg.doc_base
        push = self.ch_dense if (self.heap is not None and not self._inject_drop) else None
        if self._f8:
            # half the HBM bytes per document; q8 / q8_scale were written by the encoder's pooling kernel
            return S.sim_topk_f8(self.q8, self.q8_scale, sh.vectors_f8, sh.vec_scale, cfg.k_fetch, alive=sh.alive,
                                 id_offset=sh.cfg.doc_base, push=push, rescore=(q_emb, sh.vectors), k_fetch=32)
        return S.sim_topk(q_emb, sh.vectors, cfg.k_fetch, alive=sh.alive, id_offset=sh.cfg.doc_base, push=push)

    def _bm25_local(self):
        sh = self.shard
        # the kernel that merges the per-warp lists also pushes the shard's list to every peer (fused exchange)
        push = self.ch_bm25 if (self.heap is not None and not self._inject_drop) else None
        return sh.bm25.search(self.in_terms, k=self.cfg.k_fetch, alive=sh.alive, id_offset=sh.cfg.doc_base, push=push)

    def _exchange(self, scores, ids, chan=None):
        """per-shard top-k lists [nq, k] -> global top-k on every rank.

        p2p: the producers already pushed their lists into every rank's receive area; one merge kernel waits on the
        arrival counters and merges.  nccl: all-gather + merge (baseline)."""
        if not self.ctx.is_dist:
            return scores, ids
        if self.heap is not None:
            return S.topk_merge(chan.cand_scores, chan.cand_ids, scores.shape[1], wait=chan)
        gs = D.all_gather_cat(scores)
        gi = D.all_gather_cat(ids)
        if self.cfg.backend == "torch":
            w, nq, k = gs.shape
            flat_s = gs.permute(1, 0, 2).reshape(nq, w * k)
            flat_i = gi.permute(1, 0, 2).reshape(nq, w * k)
            v, p = torch.topk(flat_s, k, dim=1)
            return v, torch.gather(flat_i, 1, p)
        return S.topk_merge(gs, gi, scores.shape[1])

    def _fuse(self, bm_ids, de_ids):
        # (no PyTorch equivalent exists for BM25 / RRF / pair assembly: both backends use the kernels here)
        return F.rrf_fuse(bm_ids.contiguous(), de_ids.contiguous(), self.cfg.n_rerank)

    def _pairs(self, cand_ids, out_ids=None, out_lens=None):
        """(query, passage) token sequences for this rank's slice of the queries."""
        cfg, c = self.cfg, self.ctx
        q0 = c.rank * self.nq_local
        my_c = cand_ids[q0:q0 + self.nq_local].contiguous()
        pair_ids, pair_lens = F.build_pairs(self.in_q_tok[q0:q0 + self.nq_local].contiguous(),
                                            self.in_q_len[q0:q0 + self.nq_local].contiguous(), my_c, self.tok_ptrs,
                                            self.len_ptrs, self.docs_per_shard, self.passage_len, cfg.pair_seq,
                                            out_ids=out_ids, out_lens=out_lens)
        self.last_pair_lens = pair_lens          # (graph-static buffer) token count of every reranked pair
        return pair_ids, pair_lens

    def _score(self, pair_ids, pair_lens):
        """Cross-encoder logits of every rank's pairs, gathered on every rank (fp32 [nq * n_rerank])."""
        cfg, c = self.cfg, self.ctx
        if cfg.backend == "torch":
            logits = self.reranker.score_torch(pair_ids, pair_lens)
        elif cfg.varlen and pair_ids.shape[1] <= 128 and self.reranker.cfg.head_dim == 64:
            n_pairs, ch = pair_ids.shape[0], max(1, cfg.rerank_chunks)
            if ch > 1 and n_pairs % ch == 0:
                step = n_pairs // ch
                logits = torch.cat([self.reranker.score_packed(pair_ids[i:i + step], pair_lens[i:i + step].contiguous())
                                    for i in range(0, n_pairs, step)])
            else:
                logits = self.reranker.score_packed(pair_ids, pair_lens, precision=cfg.precision)
        else:
            logits = self.reranker.score(pair_ids, pair_lens)
        if self.heap is not None:
            n = logits.numel()
            self._log_stage[:n].copy_(logits.reshape(-1).float())
            return self.ch_logits(self._log_stage)[:, :n].reshape(-1)
        if c.is_dist:
            logits = D.all_gather_cat(logits.contiguous()).reshape(-1)
        return logits

    def _rerank(self, cand_ids):
        return self._score(*self._pairs(cand_ids))

    # ------------------------------------------------------------------ one batch
    def _forward(self):
        cfg = self.cfg
        if cfg.dense:
            q_emb = self._encode()
            de_s, de_i = self._dense_local(q_emb)
        bm_s, bm_i = self._bm25_local()
        if cfg.dense:
            de_s, de_i = self._exchange(de_s, de_i, getattr(self, "ch_dense", None))
        bm_s, bm_i = self._exchange(bm_s, bm_i, getattr(self, "ch_bm25", None))
        if cfg.dense:
            fu_s, fu_i = self._fuse(bm_i, de_i)
        elif cfg.rank_signals and getattr(self.shard, "neg_age", None) is not None:
            # reference search_local semantics: BM25 squash + freshness + trust + authority, fused and sorted on the device
            fu_s, fu_i = F.rank_fuse(bm_s.contiguous(), bm_i.contiguous(), cfg.n_rerank, crawled_at=self.shard.neg_age,
                                     authority=getattr(self.shard, "authority", None), now=0.0, row_base=getattr(self.shard, "signal_base", self.shard.cfg.doc_base))
        else:       # BM25 order is the fused order
            fu_s, fu_i = bm_s[:, :cfg.n_rerank].contiguous(), bm_i[:, :cfg.n_rerank].contiguous()
        if cfg.rerank:
            logits = self._rerank(fu_i)
            F.rerank_select(logits.float().contiguous(), fu_i.contiguous(), cfg.k_out, self.out_scores, self.out_ids)
        else:
            self.out_scores.copy_(fu_s[:, :cfg.k_out])
            self.out_ids.copy_(fu_i[:, :cfg.k_out])

    def load_inputs(self, enc_ids, enc_len, q_tok, q_len, terms):
        """Async H2D (or D2D) copy of one batch into the static input buffers."""
        self.in_enc_ids.copy_(enc_ids, non_blocking=True)
        self.in_enc_len.copy_(enc_len, non_blocking=True)
        self.in_q_tok.copy_(q_tok, non_blocking=True)
        self.in_q_len.copy_(q_len, non_blocking=True)
        self.in_terms.copy_(terms, non_blocking=True)

    def warm(self):
        """Capture the step graph now (index build time) instead of inside the first query."""
        if self.cfg.use_graph and not self._graph_failed and self.cfg.backend == "fused" and self._graph is None:
            self._capture()

    def run(self):
        """Execute the pipeline on the currently loaded inputs (CUDA graph replay when enabled)."""
        if self.cfg.use_graph and not self._graph_failed and self.cfg.backend == "fused":
            if self._graph is None:
                self._capture()
            if self._graph is not None:
                self._graph.replay()
                return self.out_scores, self.out_ids
        self._forward()
        return self.out_scores, self.out_ids

    def _capture(self):
        try:
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._forward()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # thread_local: other host threads (an index rebuild allocating memory, another engine replaying its graph)
            # may keep issuing CUDA calls while this thread captures
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._forward()
            self._graph = g
        except Exception as exc:  # noqa: BLE001 — fall back to eager launches, loudly
            if self.cfg.strict_graph:
                raise RuntimeError(f"CUDA graph capture failed (strict_graph): {exc!r}") from exc
            import warnings

            warnings.warn(f"CUDA graph capture failed, running eagerly: {exc!r}")
            self._graph_failed = True
            self._graph = None
            torch.cuda.synchronize()

    def search_batch(self, enc_ids, enc_len, q_tok, q_len, terms, out_scores_host=None, out_ids_host=None):
        """Public end-to-end call: pinned-host inputs -> device pipeline -> host results."""
        self.load_inputs(enc_ids, enc_len, q_tok, q_len, terms)
        self.run()
        if out_scores_host is None:
            return self.out_scores.cpu(), self.out_ids.cpu()
        out_scores_host.copy_(self.out_scores, non_blocking=True)
        out_ids_host.copy_(self.out_ids, non_blocking=True)
        return out_scores_host, out_ids_host

    # ------------------------------------------------------------------ pipelined serving (throughput mode)
    #
    # A step has two halves with opposite characters: retrieval (query encoder, BM25 + dense search, the fused top-k
    # exchange, RRF, pair assembly) is a chain of small latency-bound kernels plus one HBM stream; the cross-encoder is
    # tensor-core bound.  `submit()` overlaps them across batches with three CUDA graphs per parity on two streams:
    #
    #   stream SA:  A(i)  retrieval + pair assembly ............ then  C(i-1)  logit all-gather + final selection + D2H
    #   stream SB:  B(i)  cross-encoder on this rank's pairs (purely local compute)
    #
    # Every kernel that spins on a PEER (exchange merge, logit all-gather) lives on SA; SB never waits on anything
    # remote.  That is what makes the overlap deadlock-free: the persistent GEMM / sim_topk CTAs need (nearly) all of an
    # SM's shared memory, so a resident spinning CTA can keep one of them off its SM -- harmless as long as the spinner's
    # peer-side producer never needs this GPU's blocked kernel, which holds because every producer sits on a peer's SA
    # and SA only ever blocks on stream-level events of its own (spinner-free) SB.  Each p2p channel is used by one
    # stream only, so its use counter still advances in stream order on every rank.  NCCL exchange mode does not
    # pipeline (one communicator, two streams).
    def pipeline_supported(self) -> bool:
        cfg = self.cfg
        return (cfg.backend == "fused" and cfg.rerank and cfg.dense and cfg.use_graph and not self._graph_failed
                and (not self.ctx.is_dist or self.heap is not None))

    def _local_logits(self, pair_ids, pair_lens):
        cfg = self.cfg
        if cfg.varlen and pair_ids.shape[1] <= 128 and self.reranker.cfg.head_dim == 64:
            return self.reranker.score_packed(pair_ids, pair_lens, precision=cfg.precision)
        return self.reranker.score(pair_ids, pair_lens)

    def _pipe_init(self):
        from types import SimpleNamespace

        cfg, dev = self.cfg, self.device
        n_pairs = self.nq_local * cfg.n_rerank
        n_log = getattr(self, "_log_pad", n_pairs)
        self._pbuf = [SimpleNamespace(
            fu_s=torch.zeros((cfg.nq, cfg.n_rerank), device=dev, dtype=torch.float32),
            fu_i=torch.full((cfg.nq, cfg.n_rerank), -1, device=dev, dtype=torch.int64),
            pair_ids=torch.ones((n_pairs, cfg.pair_seq), device=dev, dtype=torch.int32),
            pair_lens=torch.ones((n_pairs,), device=dev, dtype=torch.int32),
            logits=torch.zeros((n_log,), device=dev, dtype=torch.float32),
            out_scores=torch.zeros((cfg.nq, cfg.k_out), device=dev, dtype=torch.float32),
            out_ids=torch.full((cfg.nq, cfg.k_out), -1, device=dev, dtype=torch.int64),
            host=None, done=None) for _ in range(2)]
        # SA (retrieval) outranks SB: its small latency-bound kernels and the scan CTAs take an SM as soon as one frees up
        self._sa, self._sb = torch.cuda.Stream(device=dev, priority=-1), torch.cuda.Stream(device=dev)
        from infomesh_b200 import _native

        r_sms = int(cfg.retrieval_sms)
        n_sms = _native.require().im_sm_count()
        # Both the scan and the GEMMs are persistent kernels that need an SM's whole shared memory, so they overlap only if
        # each leaves room: the scan gets r_sms CTAs, the cross-encoder GEMMs n_sms - r_sms.  Launch dimensions are frozen at
        # capture, so the budgets only have to be in force while a stage is warmed up and captured.
        budget_a = _native.sm_budgets(scan=r_sms) if r_sms > 0 else _native.sm_budgets()
        budget_b = _native.sm_budgets(gemm=n_sms - r_sms) if r_sms > 0 else _native.sm_budgets()
        cur = torch.cuda.current_stream()
        self._sa.wait_stream(cur)
        self._sb.wait_stream(cur)

        def stage_a(b):
            q_emb = self._encode()
            de_s, de_i = self._dense_local(q_emb)
            bm_s, bm_i = self._bm25_local()
            de_s, de_i = self._exchange(de_s, de_i, getattr(self, "ch_dense", None))
            bm_s, bm_i = self._exchange(bm_s, bm_i, getattr(self, "ch_bm25", None))
            F.rrf_fuse(bm_i.contiguous(), de_i.contiguous(), cfg.n_rerank, out_scores=b.fu_s, out_ids=b.fu_i)
            self._pairs(b.fu_i, out_ids=b.pair_ids, out_lens=b.pair_lens)

        def stage_b(b):
            lg = self._local_logits(b.pair_ids, b.pair_lens)
            b.logits[:n_pairs].copy_(lg.reshape(-1).float())

        def stage_c(b):
            if self.heap is not None:
                lg = self.ch_logits(b.logits)[:, :n_pairs].reshape(-1)
            else:
                lg = b.logits[:n_pairs]
            F.rerank_select(lg.contiguous(), b.fu_i, cfg.k_out, b.out_scores, b.out_ids)

        # warm every stage eagerly in pipeline order (identical on every rank), then capture one graph per stage / parity
        for b in self._pbuf:
            with torch.cuda.stream(self._sa), budget_a:
                stage_a(b)
            self._sb.wait_stream(self._sa)
            with torch.cuda.stream(self._sb), budget_b:
                stage_b(b)
            self._sa.wait_stream(self._sb)
            with torch.cuda.stream(self._sa):
                stage_c(b)
        torch.cuda.synchronize()

        def cap(fn, stream):
            out = []
            for b in self._pbuf:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                    fn(b)
                out.append(g)
            torch.cuda.synchronize()
            return out

        with budget_a:
            self._ga = cap(stage_a, self._sa)
        with budget_b:
            self._gb = cap(stage_b, self._sb)
        self._gc = cap(stage_c, self._sa)
        self._ev_a = [torch.cuda.Event() for _ in range(2)]
        self._ev_b = [torch.cuda.Event() for _ in range(2)]
        self._tick = 0
        self._pending = None      # parity whose stage C has not been enqueued yet

    def _enqueue_c(self, par):
        b = self._pbuf[par]
        with torch.cuda.stream(self._sa):
            self._sa.wait_event(self._ev_b[par])
            self._gc[par].replay()
            if b.host is not None:
                b.host[0].copy_(b.out_scores, non_blocking=True)
                b.host[1].copy_(b.out_ids, non_blocking=True)
            if b.done is not None:
                b.done.record(self._sa)

    def submit(self, enc_ids, enc_len, q_tok, q_len, terms, out_scores_host=None, out_ids_host=None, done_event=None):
        """Enqueue one batch into the pipeline and return immediately.  Its results are copied into the given (pinned)
        host tensors when its final stage runs -- one `submit` later, or at :meth:`drain`; ``done_event`` is recorded
        right after those copies.  Always finish a stream of submits with :meth:`drain`."""
        if getattr(self, "_ga", None) is None:
            self._pipe_init()
        par = self._tick & 1
        b = self._pbuf[par]
        self._tick += 1
        with torch.cuda.stream(self._sa):
            # (buffers of this parity are free: C of the batch two submits ago precedes this point on SA)
            self.load_inputs(enc_ids, enc_len, q_tok, q_len, terms)
            self._ga[par].replay()
            self._ev_a[par].record(self._sa)
        b.host = (out_scores_host, out_ids_host) if out_scores_host is not None else None
        b.done = done_event
        with torch.cuda.stream(self._sb):
            self._sb.wait_event(self._ev_a[par])
            self._gb[par].replay()
            self._ev_b[par].record(self._sb)
        if self._pending is not None:
            self._enqueue_c(self._pending)     # previous batch: all-gather + select, behind this batch's retrieval on SA
        self._pending = par

    def drain(self):
        """Finish the last batch's final stage and wait until everything submitted has completed."""
        if getattr(self, "_sa", None) is None:
            return
        if self._pending is not None:
            self._enqueue_c(self._pending)
            self._pending = None
        self._sa.synchronize()
        self._sb.synchronize()

    def pipeline_streams(self):
        if getattr(self, "_ga", None) is None:
            self._pipe_init()
        return self._sa, self._sb

    def degraded_shards(self) -> list[int]:
        """Ranks whose lists were dropped from a merge because they stayed silent (p2p exchange, ``degraded_ok``)."""
        if self.heap is None:
            return []
        return sorted(set(self.ch_dense.dead_ranks()) | set(self.ch_bm25.dead_ranks()))

    def apply_health(self, monitor) -> list[int]:
        """Feed NVML health (ECC uncorrected errors, Xid events, a lost device: ``resources/gpu_health.py``) into the
        degraded-mode mask of the exchange channels: the merge kernels read the status word first and ignore the lists of
        masked shards without waiting for them, so an unhealthy GPU costs recall on its documents, not availability.
        Only with ``degraded_ok`` and the p2p exchange; device index == local rank on a single node.  Returns the masked ranks."""
        if self.heap is None or not self.cfg.degraded_ok:
            return []
        bad = [r for r in monitor.unhealthy_devices() if 0 <= r < self.ctx.world and r != self.ctx.rank]
        if bad:
            mask = 0
            for r in bad:
                mask |= 1 << r
            for ch in (self.ch_dense, self.ch_bm25):
                ch.status |= mask                       # sticky, like a timeout; clear_status() re-admits the shard
        return bad

    def stage_times(self, iters: int = 5) -> dict:
        """Median device time (ms) of every pipeline stage, measured eagerly with CUDA events between the stages (so
        the sum exceeds a CUDA-graph replay of the whole step by the launch gaps).  Names the limiter at each N."""
        cfg = self.cfg
        names = ["encode", "dense_local", "bm25_local", "exchange_merge", "fuse_pairs", "cross_encoder", "select"]
        acc = {n: [] for n in names}
        for _ in range(iters + 1):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            ev[0].record()
            q_emb = self._encode()
            ev[1].record()
            de_s, de_i = self._dense_local(q_emb)
            ev[2].record()
            bm_s, bm_i = self._bm25_local()
            ev[3].record()
            de_s, de_i = self._exchange(de_s, de_i, getattr(self, "ch_dense", None))
            bm_s, bm_i = self._exchange(bm_s, bm_i, getattr(self, "ch_bm25", None))
            ev[4].record()
            fu_s, fu_i = self._fuse(bm_i, de_i)
            pairs = self._pairs(fu_i) if cfg.rerank else None
            ev[5].record()
            if cfg.rerank:
                logits = self._score(*pairs)
            ev[6].record()
            if cfg.rerank:
                F.rerank_select(logits.float().contiguous(), fu_i.contiguous(), cfg.k_out, self.out_scores, self.out_ids)
            ev[7].record()
            torch.cuda.synchronize()
            for j, n in enumerate(names):
                acc[n].append(ev[j].elapsed_time(ev[j + 1]))
        out = {}
        for n in names:
            v = sorted(acc[n][1:])
            out[n] = round(v[len(v) // 2], 4)
        return out

    def launches_per_step(self) -> int:
        from infomesh_b200 import _native

        _native.reset_launch_count()
        self._forward()
        torch.cuda.synchronize()
        return _native.launch_count()
