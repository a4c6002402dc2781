"""GPU numerics: every CUDA kernel against a plain PyTorch fp32 (or NumPy / pure-Python) reference of the same op."""
from __future__ import annotations

import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _native_required():
    from infomesh_b200 import _native

    _native.require()          # fail loudly if the extension is missing on a GPU box
    torch.manual_seed(0)


@pytest.mark.parametrize("m,n,k,bias,act,resid,bn,fp32", [
    (128, 128, 64, False, None, False, 0, False), (1000, 1152, 384, True, None, False, 0, False),
    (2048, 1536, 384, True, "gelu", False, 0, False), (2048, 384, 1536, True, None, True, 0, False),
    (4096, 768, 3072, True, None, True, 256, False), (300, 200, 72, False, None, False, 0, False),
    (64, 384, 384, True, "relu", False, 0, True), (4096, 3072, 768, True, "gelu", False, 256, False),
    (512, 512, 512, True, "tanh", False, 0, False),
])
def test_gemm_matches_fp32_reference(m, n, k, bias, act, resid, bn, fp32):
    from infomesh_b200.ops.gemm import linear, linear_ref

    a = (torch.randn(m, k, device=DEV) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=DEV) * 0.5).bfloat16()
    b = torch.randn(n, device=DEV) if bias else None
    r = torch.randn(m, n, device=DEV).bfloat16() if resid else None
    out = linear(a, w, b, r, act, bn=bn, out_dtype=torch.float32 if fp32 else torch.bfloat16)
    if act == "tanh":
        ref = torch.tanh(a.float() @ w.float().t() + b)
    else:
        ref = linear_ref(a, w, b, r, act)
    err = (out.float() - ref).abs().max().item()
    assert err <= 0.02 * ref.abs().max().item() + 0.05


@pytest.mark.parametrize("nq,n,dim,k,alive", [(1, 1000, 384, 10, None), (7, 5000, 384, 10, None),
                                               (64, 60000, 384, 10, None), (128, 100001, 384, 16, None),
                                               (33, 30000, 384, 32, None), (64, 50000, 384, 10, 0.5),
                                               (5, 130, 128, 10, None), (16, 20000, 512, 10, None),
                                               (64, 200000, 384, 20, 0.7), (20, 400000, 384, 20, None)])
def test_sim_topk_exact(nq, n, dim, k, alive):
    from infomesh_b200.ops.search import sim_topk, sim_topk_ref

    q = torch.nn.functional.normalize(torch.randn(nq, dim, device=DEV), dim=1).bfloat16()
    d = torch.nn.functional.normalize(torch.randn(n, dim, device=DEV), dim=1).bfloat16()
    mask = (torch.rand(n, device=DEV) < alive).to(torch.uint8) if alive is not None else None
    s, i = sim_topk(q, d, k, alive=mask, id_offset=1000)
    rs, ri = sim_topk_ref(q, d, k, alive=mask)
    kk = rs.shape[1]
    assert (s[:, :kk] - rs).abs().max().item() < 2e-3
    assert (i[:, :kk] == ri + 1000).float().mean().item() > 0.98


def test_quantize_rows_e4m3_matches_oracle():
    from infomesh_b200.ops.nn import quantize_rows_e4m3, quantize_rows_e4m3_ref

    x = (torch.randn(777, 384, device=DEV) * torch.rand(777, 1, device=DEV) * 3).bfloat16()
    x[5] = 0
    q8, sc = quantize_rows_e4m3(x)
    r8, rsc = quantize_rows_e4m3_ref(x)
    live = torch.ones(777, dtype=torch.bool, device=DEV)
    live[5] = False                                  # an all-zero row may carry any finite scale
    assert torch.equal(sc[live], rsc[live]) and torch.isfinite(sc).all() and (q8[5] == 0).all()
    deq = q8.view(torch.float8_e4m3fn).float() * sc[:, None]
    ref = r8.view(torch.float8_e4m3fn).float() * rsc[:, None]
    assert (deq - ref).abs().max().item() <= 1e-6 + 0.07 * x.float().abs().max().item()   # at most one e4m3 ulp apart
    assert (q8 != r8).float().mean().item() < 0.01                                          # ties in rounding only
    assert (deq - x.float()).abs().max().item() <= x.float().abs().amax(dim=1).max().item() / 14   # half an ulp of 3 mantissa bits


@pytest.mark.parametrize("nq,n,dim,k,alive", [(64, 300_000, 384, 20, None), (8, 50_000, 384, 10, 0.9), (128, 20_000, 512, 16, None),
                                               (1, 5000, 128, 10, None)])
def test_sim_topk_f8_scores_and_rescored_recall(nq, n, dim, k, alive):
    """fp8 shard: list scores are the dot products of the DEQUANTISED vectors; with bf16 re-scoring of a 32-wide
    over-fetch the result matches the exact bf16 search (recall@k vs the fp32 oracle >= 0.99)."""
    from infomesh_b200.ops.nn import quantize_rows_e4m3
    from infomesh_b200.ops.search import sim_topk_f8, sim_topk_ref

    q = torch.nn.functional.normalize(torch.randn(nq, dim, device=DEV), dim=1).bfloat16()
    d = torch.nn.functional.normalize(torch.randn(n, dim, device=DEV), dim=1).bfloat16()
    mask = (torch.rand(n, device=DEV) < alive).to(torch.uint8) if alive is not None else None
    q8, qs = quantize_rows_e4m3(q)
    d8, ds = quantize_rows_e4m3(d)
    # (1) raw fp8 search against the oracle over the dequantised operands
    s, i = sim_topk_f8(q8, qs, d8, ds, k, alive=mask, id_offset=7)
    qd = q8.view(torch.float8_e4m3fn).float() * qs[:, None]
    dd = d8.view(torch.float8_e4m3fn).float() * ds[:, None]
    rs, ri = sim_topk_ref(qd, dd, k, alive=mask)
    assert (s - rs).abs().max().item() < 1e-4
    assert (i == ri + 7).float().mean().item() > 0.98
    # (2) over-fetch + exact re-scoring vs the fp32 oracle on the bf16 vectors
    s2, i2 = sim_topk_f8(q8, qs, d8, ds, k, alive=mask, id_offset=7, rescore=(q, d), k_fetch=32)
    es, ei = sim_topk_ref(q, d, k, alive=mask)
    recall = sum(len(set(a.tolist()) & set((b + 7).tolist())) for a, b in zip(i2, ei)) / ei.numel()
    assert recall >= 0.99, recall
    hit = i2 == ei + 7
    assert (s2 - es).abs()[hit].max().item() < 2e-3
    assert (s2[:, :-1] >= s2[:, 1:]).all()


def test_rescore_topk_orders_and_pads():
    from infomesh_b200.ops.search import rescore_topk

    q = torch.nn.functional.normalize(torch.randn(3, 384, device=DEV), dim=1).bfloat16()
    d = torch.nn.functional.normalize(torch.randn(100, 384, device=DEV), dim=1).bfloat16()
    cand = torch.tensor([[5, 9, -1, 70, 9], [1, 2, 3, 4, 5], [-1, -1, -1, -1, -1]], device=DEV, dtype=torch.int64)
    s, i = rescore_topk(q, d, cand, 4, id_offset=100)
    full = q.float() @ d.float().t()
    for r in range(2):
        rows = [x for x in cand[r].tolist() if x >= 0]
        order = sorted(set(rows), key=lambda x: (-full[r, x].item(), x))
        got = [x - 100 for x in i[r].tolist() if x >= 0]
        assert got[0] == order[0] and set(got) <= set(rows)
        assert abs(s[r, 0].item() - full[r, order[0]].item()) < 2e-3
    assert i[2].tolist() == [-1, -1, -1, -1]


def test_topk_merge_dedup_and_order():
    from infomesh_b200.ops.search import topk_merge

    sc = torch.tensor([[[0.9, 0.5, 0.1]], [[0.9, 0.7, 0.2]]], device=DEV)                 # [P=2, nq=1, 3]
    ids = torch.tensor([[[5, 8, 2]], [[3, 8, -1]]], device=DEV, dtype=torch.int64)
    s, i = topk_merge(sc, ids, 4)
    assert i[0].tolist() == [3, 5, 8, 2] and s[0, 0].item() == pytest.approx(0.9)          # tie -> lower id; 8 deduped


@pytest.mark.parametrize("B,nH,hd,Sq,Sk,lens,causal,bias,scale,packed", [
    (2, 12, 64, 128, 128, False, False, False, None, True), (2, 12, 64, 256, 256, True, False, False, None, True),
    (3, 12, 32, 128, 128, False, False, False, None, True), (2, 12, 32, 384, 384, True, False, False, None, True),
    (2, 8, 64, 256, 256, False, True, False, None, True), (2, 8, 64, 256, 256, True, False, True, 1.0, True),
    (2, 8, 64, 128, 384, True, False, False, None, False), (2, 12, 64, 100, 100, True, False, False, None, True),
    (1, 12, 32, 40, 40, False, False, False, None, True), (5, 12, 64, 128, 128, True, False, False, None, True),
    (2, 8, 64, 128, 128, False, True, False, None, True), (3, 4, 64, 32, 32, True, False, False, None, True),
    (2, 8, 64, 64, 128, True, False, False, None, False)])
def test_attention_matches_reference(B, nH, hd, Sq, Sk, lens, causal, bias, scale, packed):
    from infomesh_b200.ops.attention import attention, attention_ref

    HH = nH * hd
    if packed and Sq == Sk:
        qkv = (torch.randn(B, Sq, 3 * HH, device=DEV) * 0.7).bfloat16()
        q, k, v = qkv[..., :HH], qkv[..., HH:2 * HH], qkv[..., 2 * HH:]
    else:
        q = (torch.randn(B, Sq, HH, device=DEV) * 0.7).bfloat16()
        kv = (torch.randn(B, Sk, 2 * HH, device=DEV) * 0.7).bfloat16()
        k, v = kv[..., :HH], kv[..., HH:]
    kv_lens = torch.randint(1, Sk + 1, (B,), device=DEV, dtype=torch.int32) if lens else None
    rb = (torch.randn(nH, Sq + Sk - 1, device=DEV) * 0.5) if bias else None
    off = Sk - Sq if causal else 0
    o = attention(q, k, v, nH, kv_lens, causal, off, scale, rb)
    ref = attention_ref(q, k, v, nH, kv_lens, causal, off, scale, rb)
    assert (o.float() - ref).abs().max().item() < 0.03


def test_attention_decode_matches_reference():
    from infomesh_b200.ops.attention import attention_decode, attention_ref

    B, nH, hd, S = 3, 8, 64, 50
    q = (torch.randn(B, nH * hd, device=DEV) * 0.7).bfloat16()
    kc = (torch.randn(B, 64, nH * hd, device=DEV) * 0.7).bfloat16()
    vc = (torch.randn(B, 64, nH * hd, device=DEV) * 0.7).bfloat16()
    lens = torch.tensor([S, 7, 64], device=DEV, dtype=torch.int32)
    o = attention_decode(q, kc, vc, nH, lens)
    ref = attention_ref(q[:, None], kc, vc, nH, kv_lens=lens)[:, 0]
    assert (o.float() - ref).abs().max().item() < 0.03


@pytest.mark.parametrize("H", [384, 768, 512])
def test_norm_embed_pool_heads(H):
    from infomesh_b200.ops import nn as N

    V, S, B = 1000, 64, 4
    ids = torch.randint(0, V, (B * S,), device=DEV, dtype=torch.int32)
    word, pos, typ = (torch.randn(n, H, device=DEV).bfloat16() for n in (V, 512, 2))
    g = torch.rand(H, device=DEV) + 0.5
    b = torch.randn(H, device=DEV)
    close = lambda x, y, tol: (x.float() - y.float()).abs().max().item() < tol  # noqa: E731
    assert close(N.embed_ln(ids, word, pos, typ, g, b, 1e-12, S), N.embed_ln_ref(ids, word, pos, typ, g, b, 1e-12, S), 0.06)
    x = torch.randn(B * S, H, device=DEV).bfloat16()
    r = torch.randn(B * S, H, device=DEV).bfloat16()
    assert close(N.layernorm(x, g, b, 1e-12, residual=r), N.layernorm_ref(x, g, b, 1e-12, r), 0.06)
    assert close(N.layernorm(x, g, None, 1e-6, rms_only=True), N.layernorm_ref(x, g, None, 1e-6, None, True), 0.06)
    parts = torch.randn(3, B * S, H, device=DEV).bfloat16()
    assert close(N.layernorm(parts, g, b, 1e-5, partials=3, partial_stride=B * S * H),
                 N.layernorm_ref(parts.float().sum(0), g, b, 1e-5), 0.08)
    h = torch.randn(B, S, H, device=DEV).bfloat16()
    lens = torch.tensor([64, 10, 33, 1], device=DEV, dtype=torch.int32)
    assert close(N.pool_norm(h, lens, "cls"), N.pool_norm_ref(h, lens, "cls"), 0.01)
    assert close(N.pool_norm(h, lens, "mean"), N.pool_norm_ref(h, lens, "mean"), 0.01)
    w1 = (torch.randn(H, H, device=DEV) / math.sqrt(H)).bfloat16()
    b1 = torch.randn(H, device=DEV) * 0.1
    w2 = (torch.randn(1, H, device=DEV) / math.sqrt(H)).bfloat16()
    b2 = torch.randn(1, device=DEV)
    ref = (torch.tanh(h[:, 0].float() @ w1.float().t() + b1) @ w2.float().t()).squeeze(1) + b2
    assert close(N.cls_head(h, w1, b1, w2, b2), ref, 0.02)
    lg = torch.randn(7, 32128, device=DEV)
    _, idx = N.row_argmax(lg, id_offset=5)
    assert (idx.long() == lg.argmax(1) + 5).all()


def test_bert_models_match_fp32_reference():
    from infomesh_b200.models.bert import BGE_RERANKER_BASE, BGE_SMALL, BertModel

    m = BertModel(BGE_SMALL, device=DEV, seed=1)
    ids = torch.randint(5, 30000, (6, 128), device=DEV, dtype=torch.int32)
    lens = torch.randint(40, 129, (6,), device=DEV, dtype=torch.int32)
    e, r = m.embed(ids, lens).float(), m.embed_ref(ids, lens)
    assert torch.nn.functional.cosine_similarity(e, r, dim=1).min().item() > 0.999
    rr = BertModel(BGE_RERANKER_BASE, device=DEV, seed=2)
    ids = torch.randint(5, 250000, (4, 128), device=DEV, dtype=torch.int32)
    lens = torch.tensor([128, 77, 30, 100], device=DEV, dtype=torch.int32)
    s, ref = rr.score(ids, lens), rr.score_ref(ids, lens)
    assert (s - ref).abs().max().item() < 0.02


def test_bm25_kernel_matches_oracle():
    from infomesh_b200.engine.synth import SynthConfig, SynthShard, make_queries
    from infomesh_b200.ops import bm25 as BM

    cfg = SynthConfig(n_docs=40_000, n_docs_global=40_000, vocab_terms=4_000, doc_len=48, passage_len=64)
    sh = SynthShard(cfg, device=DEV, build_chunk=15_000)
    csr = dict(off=sh.bm25.off.cpu().numpy(), doc=sh.bm25.doc.cpu().numpy(), tf=sh.bm25.tf.cpu().numpy(),
               doc_len=np.full(cfg.n_docs, cfg.doc_len, np.int32), df=sh.df_local.cpu().numpy())
    qt, _, _, _ = make_queries(cfg, 32, device=DEV)
    extra = torch.tensor([[5, -1, -1, -1, -1, -1, -1, -1], [3999, 3, -1, -1, -1, -1, -1, -1],
                          [7000, 1, -1, -1, -1, -1, -1, -1], [-1] * 8], dtype=torch.int32)
    qt = torch.cat([qt, extra])
    s, i = sh.bm25.search(qt.to(DEV), k=20)
    for q in range(qt.shape[0]):
        ref = BM.bm25_ref(csr, qt[q].tolist(), k=20)
        got = [(a, b) for a, b in zip(s[q].tolist(), i[q].tolist()) if b >= 0]
        assert len(got) == len(ref)
        for (rs, _), (gs, _) in zip(ref, got):
            assert abs(rs - gs) <= 1e-3 * max(rs, 1e-6)
        if ref and ref[-1][0] < ref[0][0] * 0.999:          # ids must agree wherever scores are not tied
            assert got[0][1] == ref[0][1] or abs(ref[0][0] - ref[1][0]) < 1e-6


def test_passage_simhash_hamming_fuse_kernels():
    from infomesh_b200.ops import bm25 as BM
    from infomesh_b200.ops import dedup as DD
    from infomesh_b200.ops import fuse as F

    tok = torch.randint(0, 50, (5000,), dtype=torch.int32)
    pass_off = torch.tensor(sorted(set([0, 5000] + torch.randint(1, 4999, (60,)).tolist())), dtype=torch.int64)
    n_pass = pass_off.numel() - 1
    doc_pass = torch.tensor([0, n_pass // 3, n_pass // 3, n_pass], dtype=torch.int64)
    qt = torch.tensor([[1, 2, 3, -1], [7, 9, 49, -1], [60, 61, -1, -1]], dtype=torch.int32)
    pd = torch.tensor([0, 2, 1, 2, 0], dtype=torch.int32)
    pq = torch.tensor([0, 1, 0, 2, 1], dtype=torch.int32)
    os_, op_ = BM.passage_score(tok.to(DEV), pass_off.to(DEV), doc_pass.to(DEV), pd.to(DEV), pq.to(DEV), qt.to(DEV))
    for p in range(5):
        d, q = int(pd[p]), int(pq[p])
        if doc_pass[d + 1] > doc_pass[d]:
            rs, rp = BM.passage_score_ref(tok.numpy(), pass_off[doc_pass[d]:doc_pass[d + 1] + 1].tolist(), qt[q].tolist())
        else:
            rs, rp = 0.0, -1
        assert rp == int(op_[p]) and abs(rs - float(os_[p])) < 1e-5

    texts = ["The quick brown fox jumps over the lazy dog " * 3, "", "one", "one two", "Héllo wörld ünïcode " * 5,
             "x" * 300 + " long words " + "y" * 70] + [f"doc number {i} about gpus {i * i}" for i in range(100)]
    fp = DD.simhash_batch(texts, device=DEV)
    want = np.asarray([DD.simhash_py(t) for t in texts], dtype=np.uint64).view(np.int64)
    assert (fp.cpu().numpy() == want).all()                                   # bit-exact with md5(shingle)[:8] BE
    table = torch.from_numpy(np.random.default_rng(0).integers(0, 2**63, 100_001, dtype=np.int64)).to(DEV)
    probes = table[[5, 77, 100_000]].clone()
    probes[0] ^= 0b101
    probes[1] ^= (1 << 40) | (1 << 3) | (1 << 9) | (1 << 20)
    dist, idx = DD.unpack_best(DD.hamming_scan(table, probes, 3))
    assert dist.tolist() == [2, -1, 0] and idx.tolist() == [5, -1, 100_000]

    g = torch.Generator().manual_seed(0)
    ids_a = torch.stack([torch.randperm(60, generator=g)[:20] for _ in range(64)]).long()
    ids_b = torch.stack([torch.randperm(60, generator=g)[:20] for _ in range(64)]).long()
    ids_a[3, 10:] = -1
    ids_b[4, :] = -1
    ids_a[5, :] = -1
    ids_b[5, :] = -1
    rs, ri = F.rrf_fuse_ref(ids_a, ids_b, 20)
    gs, gi = F.rrf_fuse(ids_a.to(DEV), ids_b.to(DEV), 20)
    assert (gi.cpu() == ri).all() and torch.allclose(gs.cpu()[ri >= 0], rs[ri >= 0], atol=1e-6)
    tok_store = torch.randint(1000, 5000, (60, 24), dtype=torch.int32)
    len_store = torch.randint(0, 25, (60,), dtype=torch.int32)
    q_tok = torch.randint(1000, 5000, (64, 8), dtype=torch.int32)
    q_len = torch.randint(1, 9, (64,), dtype=torch.int32)
    rp, rl = F.build_pairs_ref(q_tok, q_len, ri, tok_store, len_store, 24, 48)
    ts, ls = tok_store.to(DEV), len_store.to(DEV)
    gp, gl = F.build_pairs(q_tok.to(DEV), q_len.to(DEV), gi, F.ptr_table([ts], DEV), F.ptr_table([ls], DEV), 60, 24, 48)
    assert (gp.cpu() == rp).all() and (gl.cpu() == rl).all()
    logits = torch.randn(64, 20)
    logits[2, 3] = float("nan")
    _, si = F.rerank_select(logits.to(DEV), gi, 10)
    for q in range(64):
        lq = logits[q].clone()
        lq[(ri[q] < 0) | torch.isnan(lq)] = float("-inf")
        order = sorted(range(20), key=lambda c: (-lq[c].item(), c))[:10]
        assert [int(ri[q][c]) for c in order] == si[q].cpu().tolist()


def test_hybrid_engine_end_to_end_matches_torch_backend():
    """The fused pipeline and the PyTorch (cuBLAS/SDPA) build of the same pipeline agree on the retrieved ids."""
    from infomesh_b200.engine.hybrid import HybridConfig, HybridEngine
    from infomesh_b200.engine.synth import SynthConfig, SynthShard, make_queries
    from infomesh_b200.models.bert import BertConfig, BertModel

    cfg = SynthConfig(n_docs=30_000, n_docs_global=30_000, vocab_terms=5_000, doc_len=32, passage_len=48)
    shard = SynthShard(cfg, device=DEV)
    enc = BertModel(BertConfig(name="t-enc", layers=2), device=DEV, seed=1)
    qt, qtok, qlen, _ = make_queries(cfg, 16, device=DEV)
    enc_ids = torch.zeros((16, 32), dtype=torch.int32)
    enc_ids[:, 0] = 101
    enc_ids[:, 1:4] = (qtok[:, :3] % 20000) + 1000
    enc_ids[:, 4] = 102
    enc_len = torch.full((16,), 5, dtype=torch.int32)
    outs = {}
    for backend in ("fused", "torch"):
        eng = HybridEngine(shard, HybridConfig(nq=16, rerank=False, backend=backend, use_graph=(backend == "fused")),
                           encoder=enc)
        s, i = eng.search_batch(enc_ids.to(DEV), enc_len.to(DEV), qtok.to(DEV), qlen.to(DEV), qt.to(DEV))
        outs[backend] = i
    # position-wise equality is too strict (bf16 score ties reorder neighbours): compare the top-10 *sets*
    f, t = outs["fused"].cpu().tolist(), outs["torch"].cpu().tolist()
    overlap = sum(len(set(a) & set(b)) / max(len(set(b)), 1) for a, b in zip(f, t)) / len(f)
    assert overlap > 0.8 and (outs["fused"][:, 0] >= 0).all(), overlap


@pytest.mark.gpu
def test_symmetric_heap_single_rank_selfpush():
    """World-size-1 degenerate case of the peer-memory backend: self-push all-gather and fused top-k exchange."""
    from infomesh_b200.ops import search as S
    from infomesh_b200.parallel import symm

    dev = torch.device("cuda:0")
    heap = symm.SymmetricHeap(4 << 20)
    ag = symm.AllGatherChannel(heap, (1024,), torch.float32)
    for step in range(3):
        src = torch.randn(1024, device=dev)
        out = ag(src)
        assert torch.equal(out[0], src)
    ch = symm.TopkChannel(heap, 8, 10)
    sc = torch.rand((4, 8, 10), device=dev).sort(dim=2, descending=True).values
    ids = torch.arange(4 * 8 * 10, device=dev, dtype=torch.int64).reshape(4, 8, 10)
    for step in range(3):
        ls, li = S.topk_merge(sc, ids, 10, push=ch)
        gs, gi = S.topk_merge(ch.cand_scores, ch.cand_ids, 10, wait=ch)
        assert torch.equal(ls, gs) and torch.equal(li, gi)
    torch.cuda.synchronize()
    heap.close()


@pytest.mark.gpu
def test_t5_encoder_and_greedy_decode_match_fp32_reference():
    from infomesh_b200.models.t5 import T5_TINY, T5Model

    dev = torch.device("cuda:0")
    m = T5Model(T5_TINY, device=dev, seed=3)
    g = torch.Generator(device="cpu").manual_seed(5)
    ids = torch.randint(2, T5_TINY.vocab_size, (3, 40), generator=g, dtype=torch.int32).to(dev)
    lens = torch.tensor([40, 17, 33], dtype=torch.int32, device=dev)
    enc = m.encode(ids, lens).float()
    ref = m.encode_ref(ids, lens)
    for b in range(3):                                   # padded positions are don't-care
        L = int(lens[b])
        err = (enc[b, :L] - ref[b, :L]).abs().max().item()
        assert err < 0.08, err
    T = 6
    toks = m.generate(ids, lens, max_new_tokens=T)
    rtoks, rlogits = m.generate_ref(ids, lens, max_new_tokens=T, enc_states=ref)
    # greedy paths agree wherever the reference's top-2 margin is larger than bf16 noise
    top2 = rlogits.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1])
    same = (toks[:, :T].cpu() == rtoks.cpu())
    assert bool((same | (margin.cpu() < 0.05)).all()), (toks, rtoks, margin)


@pytest.mark.gpu
def test_gpu_search_index_over_local_store(tmp_path):
    from infomesh_b200.engine.gpu_index import GpuSearchIndex
    from infomesh_b200.index.local_store import LocalStore
    from infomesh_b200.models.bert import BertConfig, BertModel

    store = LocalStore(tmp_path / "idx.db")
    topics = ["tensor memory accumulators on blackwell", "kademlia routing table buckets", "sqlite full text search ranking",
              "merkle tree audit proofs", "simhash near duplicate detection"]
    for i in range(60):
        t = topics[i % len(topics)]
        store.add_document(url=f"https://example.org/{i}", title=f"Doc {i} {t.split()[0]}", text=f"{t} sample number {i}. " * 6,
                           raw_html_hash=f"r{i}", text_hash=f"t{i}", language="en")
    dev = torch.device("cuda:0")
    small = BertConfig(name="tiny-enc", vocab_size=30522, hidden=384, layers=2, heads=12, ffn=1536, max_pos=512)
    rr = BertConfig(name="tiny-rr", vocab_size=250002, hidden=768, layers=2, heads=12, ffn=3072, max_pos=514, pos_offset=2,
                    type_vocab=1, classifier=True)
    gi = GpuSearchIndex(store, device=dev, encoder=BertModel(small, device=dev, seed=1), reranker=BertModel(rr, device=dev, seed=2), allow_untrained=True,
                        query_batch=8)
    assert gi.rebuild() == 60
    res = gi.search_many(["kademlia buckets", "merkle audit proofs", "zzzunknownterm kademlia"], k=5)
    assert len(res) == 3 and len(res[0]) == 5 and all("url" in h and "snippet" in h for h in res[0])
    # BM25 AND semantics put the on-topic documents into the candidate set: every hit for query 0 comes from the topic
    assert sum("kademlia" in h["snippet"].lower() or "kademlia" in h["title"].lower() for h in res[0]) >= 3
    st = gi.stats()
    assert st["documents"] == 60 and st["hbm_bytes"] > 0
    # segments round-trip: a second index loads them (no re-encoding) and answers identically
    man = gi.save(tmp_path / "segments")
    assert man["n_docs"] == 60 and set(GpuSearchIndex._FILES) - {"authority.bin"} <= set(man["files"]) <= set(GpuSearchIndex._FILES)
    gi2 = GpuSearchIndex(store, device=dev, encoder=gi.encoder, reranker=gi.reranker, query_batch=8, allow_untrained=True)
    assert gi2.load(tmp_path / "segments") == 60
    again2 = gi2.search_many(["kademlia buckets", "merkle audit proofs"], k=5)
    assert [h["doc_id"] for h in again2[0]] == [h["doc_id"] for h in res[0]]
    assert [h["doc_id"] for h in again2[1]] == [h["doc_id"] for h in res[1]]
    # the ranking signals (document age at build time) are part of the segments: a loaded index fuses the same scores
    assert torch.equal(gi2.engine.shard.neg_age, gi.engine.shard.neg_age)
    (tmp_path / "segments" / "csr_tf.bin").write_bytes(b"\0" * man["files"]["csr_tf.bin"]["bytes"])
    with pytest.raises(ValueError, match="corrupt"):
        gi2.load(tmp_path / "segments")
    first = res[1][0]["doc_id"]
    assert gi.mark_deleted(first)
    again = gi.search("merkle audit proofs", k=5)
    assert all(h["doc_id"] != first for h in again)
    store.close()


@pytest.mark.gpu
def test_tensor_parallel_encoder_world1_matches_plain_model():
    """The fused TP/SP path at tp=1 (self-push through the symmetric heap, flag waits, channel advance) must equal
    the plain model bit-for-bit in structure and closely in value; two consecutive forwards exercise the use counters."""
    from dataclasses import replace

    from infomesh_b200.models.bert import BGE_RERANKER_BASE, BertModel
    from infomesh_b200.parallel.tp import TPBertModel

    cfg = replace(BGE_RERANKER_BASE, layers=3)
    B, S = 8, 128
    g = torch.Generator(device="cpu").manual_seed(3)
    ids = torch.randint(5, 5000, (B, S), generator=g, dtype=torch.int32).to(DEV)
    lens = torch.randint(30, S + 1, (B,), generator=g, dtype=torch.int32).to(DEV)
    ref = BertModel(cfg, device=DEV, seed=5)
    tpm = TPBertModel(cfg, B, S, seed=5, comm="fused")
    want = ref.score(ids, lens)
    for _ in range(2):
        got = tpm.score(ids, lens)
        assert (got - want).abs().max().item() < 0.03
    torch.cuda.synchronize()
    tpm.heap.close()


def test_packed_cross_encoder_matches_padded():
    """Unpadded (varlen) forward == padded forward on the valid tokens: logits agree, pad rows never computed."""
    from dataclasses import replace

    from infomesh_b200.models.bert import BGE_RERANKER_BASE, BertModel

    cfg = replace(BGE_RERANKER_BASE, layers=3)
    m = BertModel(cfg, device=DEV, seed=5)
    B, S = 37, 128
    g = torch.Generator(device="cpu").manual_seed(3)
    ids = torch.randint(5, 5000, (B, S), generator=g, dtype=torch.int32).to(DEV)
    lens = torch.randint(1, S + 1, (B,), generator=g, dtype=torch.int32).to(DEV)
    lens[0], lens[-1] = S, 1
    a = m.score(ids, lens)
    b = m.score_packed(ids, lens)
    assert torch.isfinite(b).all()
    assert (a - b).abs().max().item() < 0.03
    # garbage (even NaN) behind the packed rows must not leak into valid outputs
    x, cu, total = m.hidden_states_packed(ids, lens)
    assert int(total.item()) == int(lens.sum().item()) and cu[-1].item() == total.item()
    ref = m.hidden_states(ids, lens)
    for bi in (0, 5, B - 1):
        n = int(lens[bi].item())
        s0 = int(cu[bi].item())
        assert (x[s0:s0 + n].float() - ref[bi, :n].float()).abs().max().item() < 0.12


def test_pipelined_engine_matches_single_batch_path():
    """submit()/drain() (two batches in flight on two streams, 3 graphs per parity) returns exactly what run() returns,
    batch for batch, and the smoke() configuration of the driver runs through the unpadded cross-encoder."""
    import __graft_entry__ as entry
    from infomesh_b200.engine.hybrid import HybridConfig, HybridEngine
    from infomesh_b200.engine.synth import SynthConfig, SynthShard, make_queries
    from infomesh_b200.models.bert import BertConfig, BertModel

    entry.smoke()
    cfg = SynthConfig(n_docs=30_000, n_docs_global=30_000, vocab_terms=5_000, doc_len=32, passage_len=48)
    shard = SynthShard(cfg, device=DEV)
    enc = BertModel(BertConfig(name="t-enc", layers=2), device=DEV, seed=1)
    rr = BertModel(BertConfig(name="t-rr", vocab_size=250002, hidden=768, layers=2, heads=12, ffn=3072, max_pos=514,
                              type_vocab=1, eps=1e-5, pos_offset=2, classifier=True), device=DEV, seed=2)
    nq, nb = 8, 5
    eng = HybridEngine(shard, HybridConfig(nq=nq, pair_seq=64, use_graph=True), encoder=enc, reranker=rr)
    assert eng.pipeline_supported()
    qt, qtok, qlen, _ = make_queries(cfg, nq * nb, device=DEV)
    enc_ids = torch.zeros((nq * nb, 32), dtype=torch.int32, device=DEV)
    enc_ids[:, 0] = 101
    enc_ids[:, 1:4] = (qtok[:, :3].to(DEV) % 20000) + 1000
    enc_ids[:, 4] = 102
    enc_len = torch.full((nq * nb,), 5, dtype=torch.int32, device=DEV)
    batches = [tuple(x[i * nq:(i + 1) * nq].to(DEV).contiguous() for x in (enc_ids, enc_len, qtok, qlen, qt)) for i in range(nb)]
    ref = []
    for b in batches:
        s, i = eng.search_batch(*b)
        ref.append((s.clone(), i.clone()))
    outs = [(torch.empty((nq, 10), dtype=torch.float32).pin_memory(), torch.empty((nq, 10), dtype=torch.int64).pin_memory())
            for _ in range(nb)]
    for b, (hs, hi) in zip(batches, outs):
        eng.submit(*b, out_scores_host=hs, out_ids_host=hi)
    eng.drain()
    for (rs, ri), (hs, hi) in zip(ref, outs):
        assert torch.equal(ri.cpu(), hi)
        assert (rs.cpu() - hs).abs().max().item() < 1e-3


@pytest.mark.parametrize("m,n,k,act,res", [(300, 512, 768, None, False), (1000, 768, 3072, "gelu", False),
                                            (257, 2304, 704, None, True), (64, 256, 128, "relu", False)])
def test_gemm_fp8_matches_dequantised_reference(m, n, k, act, res):
    """e4m3 operands (per-token activation scales, per-tensor weight scale) through kind::f8f6f4 == fp32 GEMM of the
    dequantised operands; the quantiser itself stays within e4m3 rounding of the input."""
    from infomesh_b200.ops import gemm as G

    a = (torch.randn(m, k, device=DEV) * 0.7).bfloat16()
    w = (torch.randn(n, k, device=DEV) * 0.05).bfloat16()
    b = torch.randn(n, device=DEV)
    r = torch.randn(m, n, device=DEV).bfloat16() if res else None
    a8, rs = G.quantize_rows_fp8(a)
    w8, ws = G.quantize_weight_fp8(w)
    deq = a8.view(torch.float8_e4m3fn).float() * rs[:, None]
    assert ((deq - a.float()).abs() <= 0.0701 * a.float().abs() + 2e-3 * rs[:, None] * 448).all()   # e4m3: 3 mantissa bits
    out = G.linear(a8, w8, bias=b, residual=r, act=act, alpha=ws, row_scale=rs)
    ref = G.linear_fp8_ref(a8, rs, w8, ws, b, r, act)
    err = (out.float() - ref).abs().max().item()
    assert err <= 0.02 * ref.abs().max().item() + 0.05, err


def test_cross_encoder_fp8_tracks_bf16():
    from dataclasses import replace

    from infomesh_b200.models.bert import BGE_RERANKER_BASE, BertModel

    m = BertModel(replace(BGE_RERANKER_BASE, layers=3), device=DEV, seed=5)
    B, S = 24, 128
    g = torch.Generator(device="cpu").manual_seed(4)
    ids = torch.randint(5, 5000, (B, S), generator=g, dtype=torch.int32).to(DEV)
    lens = torch.randint(20, S + 1, (B,), generator=g, dtype=torch.int32).to(DEV)
    a = m.score_packed(ids, lens)
    b = m.score_packed(ids, lens, precision="fp8")
    assert torch.isfinite(b).all()
    assert (a - b).abs().max().item() < 0.25 * max(1.0, a.abs().max().item())


def test_pagerank_matches_numpy_power_iteration():
    import numpy as np

    from infomesh_b200.ops.graph import pagerank, pagerank_ref

    rng = np.random.default_rng(0)
    n, e = 5000, 60000
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    w = rng.integers(1, 5, e).astype(np.float64)
    got = pagerank(src, dst, w, n).cpu().numpy()
    ref = pagerank_ref(src, dst, w, n)
    assert np.abs(got - ref).max() < 1e-6 and abs(got.sum() - ref.sum()) < 1e-3


def test_index_builder_dedups_within_and_across_batches():
    """encode + SimHash + fingerprint scan pipeline (world 1): near-duplicates are dropped whether the original sits in
    the same batch or in the index already; the keep-mask equals the Python SimHash oracle's."""
    import random

    from infomesh_b200.engine.index_build import IndexBuilder
    from infomesh_b200.models.bert import BertConfig, BertModel
    from infomesh_b200.ops import dedup as DD

    rng = random.Random(0)
    vocab = [f"w{i}" for i in range(3000)]
    base = [" ".join(rng.choices(vocab, k=80)) for _ in range(96)]
    b1 = base[:64] + [base[3], base[10] + " extra", base[20]]                 # 3 in-batch (near-)duplicates
    b2 = base[64:] + [base[5], base[70]] + [" ".join(rng.choices(vocab, k=80)) for _ in range(30)]   # dup of indexed + in-batch
    enc = BertModel(BertConfig(name="t-enc", layers=1), device=DEV, seed=1)
    kept_fp: list[int] = []

    def oracle(texts):
        mask = []
        for t in texts:
            f = DD.simhash_py(t)
            dup = any(DD.hamming(f, g) <= DD.HAMMING_THRESHOLD for g in kept_fp)
            mask.append(not dup)
            if not dup:
                kept_fp.append(f)
        return mask

    first = 0
    ib = None
    for texts in (b1, b2):
        n = len(texts)
        if ib is None:
            ib = IndexBuilder(1000, n, encoder=enc, device=DEV)
        text, ws, we, off = DD.normalize_batch(texts)
        arrs = [torch.from_numpy(x).to(DEV) for x in (text, ws, we, off)]
        ids = torch.randint(1000, 20000, (n, 32), dtype=torch.int32, device=DEV)
        lens = torch.full((n,), 32, dtype=torch.int32, device=DEV)
        mask = ib.add_batch(ids, lens, *arrs, first_doc_id=first)
        assert mask.cpu().tolist() == oracle(texts)
        first += n
    st = ib.stats()
    assert st["indexed"] == len(kept_fp) and st["duplicates"] == st["seen"] - st["indexed"] and st["duplicates"] >= 4


def test_dedup_resolve_and_device_fill_level():
    """In-batch all-pairs dedup + min over ranks' scan results in one kernel, against the NumPy oracle; the Hamming scan
    honours a device-side table length."""
    from infomesh_b200.ops import dedup as DD

    g = torch.Generator().manual_seed(5)
    n = 700
    fp = torch.randint(-2 ** 62, 2 ** 62, (n,), generator=g, dtype=torch.int64)
    fp[100] = fp[7] ^ 0b101          # 2 bits away from an earlier passage -> dropped
    fp[650] = fp[300]                # exact repeat
    fp[5] = fp[600] ^ 1              # near a LATER passage: the later one is the duplicate
    best = torch.full((3, n), -1, dtype=torch.int64)
    best[1, 42] = (2 << 32) | 17     # rank 1 found passage 42 in its shard
    fp_d, best_d = fp.to(DEV), best.to(DEV)
    keep = DD.dedup_resolve(fp_d, best_d)
    ref = DD.dedup_resolve_ref(fp_d, best_d)
    assert keep.cpu().numpy().tolist() == ref.tolist()
    assert keep[100] == 0 and keep[650] == 0 and keep[42] == 0 and keep[5] == 1 and keep[600] == 0 and int(keep.sum()) == n - 4
    table = torch.randint(-2 ** 62, 2 ** 62, (5000,), generator=g, dtype=torch.int64).to(DEV)
    probes = table[[10, 4000]].clone()
    full = DD.unpack_best(DD.hamming_scan(table, probes))[1].tolist()
    part = DD.unpack_best(DD.hamming_scan(table, probes, n_table_dev=torch.tensor([1000], device=DEV)))[1].tolist()
    assert full == [10, 4000] and part == [10, -1]


def test_rank_fuse_matches_cpu_ranking():
    """K12: the device six-signal fuse orders candidates exactly like index.ranking.rank_results."""
    from infomesh_b200.index import ranking as R
    from infomesh_b200.ops.fuse import rank_fuse

    g = torch.Generator().manual_seed(2)
    nq, n, n_docs = 5, 20, 300
    rows = torch.stack([torch.randperm(n_docs, generator=g)[:n] for _ in range(nq)]).to(torch.int64)
    rows[1, 15:] = -1
    bm = torch.rand(nq, n, generator=g) * 12
    age = torch.rand(n_docs, generator=g) * 30 * 86400
    trust = torch.rand(n_docs, generator=g)
    auth = torch.rand(n_docs, generator=g)
    title = torch.rand(nq, n, generator=g)
    url = torch.rand(nq, n, generator=g)
    s, r, sig = rank_fuse(bm.to(DEV), rows.to(DEV), 10, crawled_at=(-age).to(DEV), trust=trust.to(DEV), authority=auth.to(DEV),
                          title_match=title.to(DEV), url_path=url.to(DEV), now=0.0, want_signals=True)
    for q in range(nq):
        cands = [R.RawCandidate(doc_id=int(rows[q, j]), url="", title="", snippet="", bm25_raw=float(bm[q, j]), crawled_at=-float(age[rows[q, j]]),
                                trust=float(trust[rows[q, j]]), authority=float(auth[rows[q, j]]), title_match=float(title[q, j]),
                                url_path=float(url[q, j])) for j in range(n) if rows[q, j] >= 0]
        want = R.rank_results(cands, limit=10, now=1e-9)
        got_ids = [int(x) for x in r[q].cpu().tolist() if x >= 0]
        assert got_ids == [w.doc_id for w in want]
        assert abs(float(s[q, 0]) - want[0].combined_score) < 2e-4
        assert abs(float(sig[q, 0, 1]) - want[0].freshness_score) < 2e-4


def test_bm25_dense_term_maps_match_plain_path_and_oracle():
    """Queries that contain very common terms: the direct tf maps of dense terms give exactly the plain (binary-search)
    result and the NumPy/FTS5-formula oracle's ordering, including two-common-term queries whose driving list is long."""
    from infomesh_b200.engine.synth import SynthConfig, SynthShard, make_queries
    from infomesh_b200.ops.bm25 import bm25_ref

    cfg = SynthConfig(n_docs=200_000, n_docs_global=200_000, vocab_terms=20_000, doc_len=48)
    shard = SynthShard(cfg, device=DEV)
    bm = shard.bm25
    n_dense = int((bm.dense_slot >= 0).sum().item()) if bm.dense_slot is not None else 0
    assert n_dense >= 5                                             # Zipf head terms cover > 1/32 of the corpus
    qt, _, _, _ = make_queries(cfg, 24, mix="common")
    qt = qt.to(DEV)
    qt[0, :3] = torch.tensor([0, 1, -1], dtype=torch.int32)          # the two MOST common terms: a ~100k-entry driving list
    qt[1, :3] = torch.tensor([2, 5, 7], dtype=torch.int32)
    s_dense, i_dense = bm.search(qt, k=20)
    keep = (bm.dense_slot, bm.dense_tf)
    bm.dense_slot = bm.dense_tf = None
    s_plain, i_plain = bm.search(qt, k=20)
    bm.dense_slot, bm.dense_tf = keep
    assert torch.equal(i_dense, i_plain) and torch.allclose(s_dense, s_plain)
    csr = {"off": bm.off.cpu().numpy(), "doc": bm.doc.cpu().numpy(), "tf": bm.tf.cpu().numpy(),
           "doc_len": np.full(cfg.n_docs, cfg.doc_len, np.int32), "df": (bm.off[1:] - bm.off[:-1]).cpu().numpy().astype(np.int32)}
    for q in (0, 1, 5):
        ref = bm25_ref(csr, qt[q].cpu().numpy(), k=20)
        got = [(round(float(s), 3), int(i)) for s, i in zip(s_dense[q].tolist(), i_dense[q].tolist()) if i >= 0]
        assert [i for _, i in got] == [i for _, i in ref] or [s for s, _ in got] == [round(s, 3) for s, _ in ref]
