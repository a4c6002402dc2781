"""Synthetic corpus + query generator, built directly on the GPU.

Documents are Zipfian term-id streams (fixed seed).  From them we derive, per shard:
  * CSR postings (term -> sorted doc ids + tf) for the BM25 kernel,
  * reranker passage tokens (term id -> model token id),
  * random unit vectors standing in for passage embeddings (the search benchmark measures the index, the
    index-build benchmark measures the encoder), and
  * queries made of 2-3 of the rarer terms of a random document so the implicit-AND BM25 query has matches.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from infomesh_b200.ops.bm25 import B as BM25_B
from infomesh_b200.ops.bm25 import K1 as BM25_K1
from infomesh_b200.ops.bm25 import Bm25Index


@dataclass
class SynthConfig:
    n_docs: int = 100_000          # documents in THIS shard
    n_docs_global: int = 100_000
    doc_base: int = 0              # global id of the shard's first document
    vocab_terms: int = 200_000
    doc_len: int = 64
    passage_len: int = 96
    dim: int = 384
    zipf_s: float = 1.0
    seed: int = 1234
    model_vocab: int = 250_002
    first_token: int = 1000


def zipf_cdf(V: int, s: float, device) -> torch.Tensor:
    """Zipf CDF, always computed on the CPU (bit-identical everywhere) and then moved."""
    w = 1.0 / torch.arange(1, V + 1, dtype=torch.float64).pow(s)
    return (w.cumsum(0) / w.sum()).to(torch.float32).to(device)


_M1 = 6364136223846793005
_M2 = -4417276706812531889   # 0xC2B2AE3D27D4EB4F as int64
_INC = 1442695040888963407


def hash_uniform(idx: torch.Tensor, seed: int) -> torch.Tensor:
    """Counter-based uniform [0,1) floats from int64 indices: integer-only mixing, so the stream is identical
    on CPU and CUDA and independent of how the corpus is sharded or chunked."""
    x = idx * _M1 + ((seed * _INC) & 0x7FFFFFFFFFFFFFFF)
    x = x ^ ((x >> 33) & 0x7FFFFFFF)
    x = x * _M2
    x = x ^ ((x >> 29) & 0x7FFFFFFFF)
    x = x * _M1
    x = x ^ ((x >> 32) & 0xFFFFFFFF)
    return ((x >> 40) & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0)


def gen_doc_terms(cfg: SynthConfig, device, start: int, count: int, cdf: torch.Tensor) -> torch.Tensor:
    """Term ids [count, doc_len] of the documents with GLOBAL ids [doc_base + start, +count)."""
    out = torch.empty((count, cfg.doc_len), device=device, dtype=torch.int32)
    col = torch.arange(cfg.doc_len, device=device, dtype=torch.int64)[None, :]
    chunk = 1 << 18
    for a in range(0, count, chunk):
        b = min(count, a + chunk)
        rows = torch.arange(cfg.doc_base + start + a, cfg.doc_base + start + b, device=device, dtype=torch.int64)
        u = hash_uniform(rows[:, None] * cfg.doc_len + col, cfg.seed)
        out[a:b] = torch.searchsorted(cdf, u).to(torch.int32).clamp_(max=cfg.vocab_terms - 1)
    return out


def gen_vectors(cfg: SynthConfig, device, start: int, count: int) -> torch.Tensor:
    """Unit-norm bf16 vectors of global documents [doc_base + start, +count) (Box-Muller on hashed uniforms)."""
    rows = torch.arange(cfg.doc_base + start, cfg.doc_base + start + count, device=device, dtype=torch.int64)
    col = torch.arange(cfg.dim, device=device, dtype=torch.int64)[None, :]
    idx = rows[:, None] * cfg.dim + col
    u1 = hash_uniform(idx, cfg.seed + 17).clamp_(min=1e-7)
    u2 = hash_uniform(idx, cfg.seed + 31)
    v = torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(6.283185307179586 * u2)
    return torch.nn.functional.normalize(v, dim=1).to(torch.bfloat16)


def term_to_token(terms: torch.Tensor, cfg: SynthConfig) -> torch.Tensor:
    span = cfg.model_vocab - cfg.first_token
    return (cfg.first_token + (terms.long() * 2654435761 % span)).to(torch.int32)


def build_csr(terms: torch.Tensor, vocab: int):
    """terms [n, L] int32 -> (off int64[V+1], doc int32[nnz], tf uint8[nnz], df int64[V]) on the same device."""
    n, L = terms.shape
    dev = terms.device
    docs = torch.arange(n, device=dev, dtype=torch.int64).repeat_interleave(L)
    key = terms.reshape(-1).long() * n + docs
    del docs
    uniq, counts = torch.unique(key, sorted=True, return_counts=True)
    del key
    t = torch.div(uniq, n, rounding_mode="floor")
    d = (uniq - t * n).to(torch.int32)
    del uniq
    df = torch.bincount(t, minlength=vocab)
    off = torch.zeros(vocab + 1, device=dev, dtype=torch.int64)
    off[1:] = df.cumsum(0)
    tf = counts.clamp_(max=255).to(torch.uint8)
    return off, d, tf, df


class SynthShard:
    """Everything one rank holds for the synthetic benchmark."""

    def __init__(self, cfg: SynthConfig, device="cuda", df_allreduce=None, build_chunk: int = 2_000_000,
                 passages: str = "local", alloc=None):
        """``passages``: "local" keeps passage tokens of this shard's documents only (multi-GPU engines read peers'
        passages through pointer tables into the symmetric heap); "global" replicates the token store of every document
        (round-1 layout, kept for A/B).  ``alloc(shape, dtype) -> tensor`` places the passage arrays (e.g. in a symmetric
        heap so peers can map them); default ``torch.empty`` on ``device``."""
        self.cfg = cfg
        dev = torch.device(device)
        self.device = dev
        cdf = zipf_cdf(cfg.vocab_terms, cfg.zipf_s, dev)
        self.cdf = cdf
        n = cfg.n_docs
        # ---- postings, built chunk-wise (chunks are contiguous doc ranges so per-term lists stay sorted) ----
        offs, docs_l, tfs_l = [], [], []
        df = torch.zeros(cfg.vocab_terms, device=dev, dtype=torch.int64)
        n_pass = n if passages == "local" else cfg.n_docs_global
        self.passages_global = passages != "local"
        mk = alloc or (lambda shape, dtype: torch.empty(shape, device=dev, dtype=dtype))
        self.passage_tok = mk((n_pass, cfg.passage_len), torch.int32)
        self.passage_len = mk((n_pass,), torch.int32)
        self.passage_len.fill_(min(cfg.passage_len, cfg.doc_len))
        p_off = cfg.doc_base if self.passages_global else 0
        for a in range(0, n, build_chunk):
            b = min(n, a + build_chunk)
            terms = gen_doc_terms(cfg, dev, a, b - a, cdf)
            pl = min(cfg.passage_len, cfg.doc_len)
            self.passage_tok[p_off + a:p_off + b, :pl] = term_to_token(terms[:, :pl], cfg)
            if pl < cfg.passage_len:
                self.passage_tok[p_off + a:p_off + b, pl:] = 1
            off, d, tf, dfc = build_csr(terms, cfg.vocab_terms)
            offs.append(off)
            docs_l.append(d + a)
            tfs_l.append(tf)
            df += dfc
            del terms
        if len(offs) == 1:
            off, doc, tf = offs[0], docs_l[0], tfs_l[0]
        else:
            # interleave chunk segments term by term: position = global term offset + offset inside earlier chunks
            off = torch.zeros(cfg.vocab_terms + 1, device=dev, dtype=torch.int64)
            off[1:] = df.cumsum(0)
            doc = torch.empty(int(off[-1].item()), device=dev, dtype=torch.int32)
            tf = torch.empty_like(doc, dtype=torch.uint8)
            run = off[:-1].clone()
            for o, d, t in zip(offs, docs_l, tfs_l):
                lens = o[1:] - o[:-1]
                term_of = torch.repeat_interleave(torch.arange(cfg.vocab_terms, device=dev), lens)
                within = torch.arange(d.numel(), device=dev) - o[:-1][term_of]
                pos = run[term_of] + within
                doc[pos] = d
                tf[pos] = t
                run += lens
                del term_of, within, pos
        del offs, docs_l, tfs_l
        if self.passages_global:
            pl = min(cfg.passage_len, cfg.doc_len)
            for a in range(0, cfg.n_docs_global, build_chunk):
                b = min(cfg.n_docs_global, a + build_chunk)
                if a >= cfg.doc_base and b <= cfg.doc_base + n:
                    continue  # already filled from the local pass
                terms = gen_doc_terms(cfg, dev, a - cfg.doc_base, b - a, cdf)
                self.passage_tok[a:b, :pl] = term_to_token(terms[:, :pl], cfg)
                if pl < cfg.passage_len:
                    self.passage_tok[a:b, pl:] = 1
                del terms
        self.df_local = df
        df_global = df.clone()
        if df_allreduce is not None:
            df_allreduce(df_global)
        N = cfg.n_docs_global
        idf = torch.log((N - df_global.double() + 0.5) / (df_global.double() + 0.5)).clamp_(min=1e-6).float()
        avg_len = float(cfg.doc_len)
        norm = torch.full((n,), BM25_K1 * (1 - BM25_B + BM25_B * cfg.doc_len / avg_len), device=dev, dtype=torch.float32)
        self.bm25 = Bm25Index.from_device_csr(off, doc, tf, norm, idf, avg_len)
        self.df_global = df_global
        # ---- dense vectors ----
        self.vectors = torch.empty((n, cfg.dim), device=dev, dtype=torch.bfloat16)
        for a in range(0, n, 1 << 20):
            b = min(n, a + (1 << 20))
            self.vectors[a:b] = gen_vectors(cfg, dev, a, b - a)
        self.alive = None

    def nbytes(self) -> int:
        return (self.vectors.numel() * 2 + self.bm25.nbytes() + self.passage_tok.numel() * 4 + self.passage_len.numel() * 4)


def make_queries(cfg: SynthConfig, n_queries: int, max_terms: int = 8, max_q_tokens: int = 32, seed: int = 99,
                 df_global: torch.Tensor | None = None, device="cpu", n_terms=(2, 3), mix: str = "rare"):
    """Queries as (term ids [nq, max_terms] -1 padded, model tokens [nq, max_q_tokens], token lens [nq]).

    Each query takes 2-3 terms of a random (global) document so AND-BM25 has at least one hit.  ``mix="rare"``: the
    terms sit in the 45-90 % rarity quantiles of that document (short posting lists); ``mix="common"``: every other
    query additionally carries the document's MOST frequent term (posting lists of ~10-40 % of the corpus), which is
    what stresses list intersection.
    Generated on ``device`` deterministically; identical on every rank.
    """
    dev = torch.device(device)
    g = torch.Generator(device="cpu").manual_seed(seed)
    doc_ids = torch.randint(0, cfg.n_docs_global, (n_queries,), generator=g)
    k_terms = torch.randint(n_terms[0], n_terms[1] + 1, (n_queries,), generator=g)
    cdf = zipf_cdf(cfg.vocab_terms, cfg.zipf_s, dev)
    q_terms = torch.full((n_queries, max_terms), -1, dtype=torch.int32)
    q_tok = torch.full((n_queries, max_q_tokens), 1, dtype=torch.int32)
    q_len = torch.zeros((n_queries,), dtype=torch.int32)
    col = torch.arange(cfg.doc_len, device=dev, dtype=torch.int64)
    for i in range(n_queries):
        d = int(doc_ids[i])
        # regenerate that document's terms exactly as the owning shard did (counter-based stream)
        u = hash_uniform(d * cfg.doc_len + col, cfg.seed)
        row = torch.searchsorted(cdf, u).clamp_(max=cfg.vocab_terms - 1)
        terms = torch.unique(row).cpu().sort().values  # ascending id == common -> rare under Zipf
        kt = min(int(k_terms[i]), terms.numel())
        # spread the picks over the 45%..90% rarity quantiles: selective but not unique terms
        pos = [int(round((terms.numel() - 1) * (0.45 + 0.45 * j / max(kt - 1, 1)))) for j in range(kt)]
        if mix == "common" and i % 2 == 1:
            pos[0] = 0
        pick = terms[torch.tensor(sorted(set(pos)), dtype=torch.long)]
        q_terms[i, :pick.numel()] = pick.to(torch.int32)
        toks = term_to_token(pick.to(torch.int32), cfg)
        q_tok[i, :toks.numel()] = toks
        q_len[i] = toks.numel()
    return q_terms, q_tok, q_len, doc_ids
