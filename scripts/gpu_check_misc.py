"""GPU checks: BM25 / passage / SimHash / Hamming / RRF / pair assembly / rerank select vs CPU oracles."""
import sys, numpy as np, torch
from infomesh_b200.ops import bm25 as BM, dedup as DD, fuse as F
from infomesh_b200.engine.synth import SynthConfig, SynthShard, make_queries
dev = "cuda"; ok_all = True
def report(name, ok, **kw):
    global ok_all
    ok_all &= bool(ok); print(dict(test=name, ok=bool(ok), **kw), flush=True)

# ---- BM25 vs oracle on a synthetic shard
cfg = SynthConfig(n_docs=50_000, n_docs_global=50_000, vocab_terms=5_000, doc_len=48, passage_len=64)
sh = SynthShard(cfg, device=dev, build_chunk=20_000)      # multi-chunk merge path
csr = dict(off=sh.bm25.off.cpu().numpy(), doc=sh.bm25.doc.cpu().numpy(), tf=sh.bm25.tf.cpu().numpy(),
           doc_len=np.full(cfg.n_docs, cfg.doc_len, np.int32), df=sh.df_local.cpu().numpy())
# postings sorted per term?
off = csr["off"]; bad = 0
for t in range(0, 5000, 97):
    d = csr["doc"][off[t]:off[t+1]]
    bad += int((np.diff(d) <= 0).sum())
report("csr_sorted", bad == 0, bad=bad)
qt, qtok, qlen, dids = make_queries(cfg, 64, device=dev)
s, i = sh.bm25.search(qt.to(dev), k=20)
torch.cuda.synchronize()
mism = 0; maxerr = 0.0
for q in range(64):
    ref = BM.bm25_ref(csr, qt[q].tolist(), k=20)
    got = [(float(a), int(b)) for a, b in zip(s[q].tolist(), i[q].tolist()) if b >= 0]
    if [d for _, d in ref] != [d for _, d in got]:
        # allow order swaps on near-equal scores
        if sorted(d for _, d in ref) != sorted(d for _, d in got): mism += 1
    for (rs, _), (gs, _) in zip(ref, got): maxerr = max(maxerr, abs(rs - gs) / max(rs, 1e-9))
report("bm25_vs_ref", mism == 0 and maxerr < 1e-3, mismatched_queries=mism, max_rel_err=maxerr)
# single-term + missing-term queries
q2 = torch.tensor([[5, -1, -1, -1], [4999, 3, -1, -1], [7000, 1, -1, -1], [-1, -1, -1, -1]], dtype=torch.int32, device=dev)
s2, i2 = sh.bm25.search(q2, k=10); torch.cuda.synchronize()
ok = True
for q in range(4):
    ref = BM.bm25_ref(csr, q2[q].tolist(), k=10)
    got = [int(b) for b in i2[q].tolist() if b >= 0]
    ok &= sorted(d for _, d in ref) == sorted(got)
report("bm25_edge", ok)

# ---- passage score
tok = torch.randint(0, 50, (5000,), dtype=torch.int32)
pass_off = torch.tensor(sorted(set([0, 5000] + torch.randint(1, 4999, (60,)).tolist())), dtype=torch.int64)
n_pass = pass_off.numel() - 1
doc_pass = torch.tensor([0, n_pass // 3, n_pass // 3, n_pass], dtype=torch.int64)   # doc1 has zero passages
qterms = torch.tensor([[1, 2, 3, -1], [7, 7, 9, 49], [60, 61, -1, -1]], dtype=torch.int32)
pd = torch.tensor([0, 2, 1, 2, 0], dtype=torch.int32); pq = torch.tensor([0, 1, 0, 2, 1], dtype=torch.int32)
qt_u = qterms.clone(); qt_u[1] = torch.tensor([7, 9, 49, -1])       # kernel expects unique terms
os_, op_ = BM.passage_score(tok.to(dev), pass_off.to(dev), doc_pass.to(dev), pd.to(dev), pq.to(dev), qt_u.to(dev))
torch.cuda.synchronize(); ok = True
for p in range(5):
    d, q = int(pd[p]), int(pq[p])
    bounds = pass_off[doc_pass[d]:doc_pass[d + 1] + 1].tolist() if doc_pass[d + 1] > doc_pass[d] else [0]
    rs, rp = BM.passage_score_ref(tok.numpy(), bounds, qt_u[q].tolist()) if len(bounds) > 1 else (0.0, -1)
    ok &= (rp == int(op_[p])) and abs(rs - float(os_[p])) < 1e-5
report("passage_score", ok, got=op_.tolist())

# ---- simhash + hamming
texts = ["The quick brown fox jumps over the lazy dog " * 3, "", "one", "one two", "Héllo wörld ünïcode test test " * 5,
         "a b c d e f g " * 40, "x" * 300 + " long words " + "y" * 70, "Python asyncio tutorial"] + [f"doc number {i} about search engines and gpus {i*i}" for i in range(200)]
fp = DD.simhash_batch(texts, device=dev); torch.cuda.synchronize()
want = np.asarray([DD.simhash_py(t) for t in texts], dtype=np.uint64).view(np.int64)
report("simhash_bit_exact", bool((fp.cpu().numpy() == want).all()), n=len(texts))
table = torch.from_numpy(np.random.default_rng(0).integers(0, 2**63, 100_001, dtype=np.int64)).to(dev)
probes = table[[5, 77, 100_000]].clone(); probes[0] ^= 0b101; probes[1] ^= (1 << 40) | (1 << 3) | (1 << 9) | (1 << 20)
best = DD.hamming_scan(table, probes, 3); dist, idx = DD.unpack_best(best); torch.cuda.synchronize()
report("hamming_scan", dist.tolist() == [2, -1, 0] and idx.tolist() == [5, -1, 100_000], dist=dist.tolist(), idx=idx.tolist())

# ---- rrf / pairs / select
g = torch.Generator().manual_seed(0)
ids_a = torch.stack([torch.randperm(60, generator=g)[:20] for _ in range(64)]).long()
ids_b = torch.stack([torch.randperm(60, generator=g)[:20] for _ in range(64)]).long()
ids_a[3, 10:] = -1; ids_b[4, :] = -1; ids_a[5, :] = -1; ids_b[5, :] = -1
rs, ri = F.rrf_fuse_ref(ids_a, ids_b, 20)
gs, gi = F.rrf_fuse(ids_a.to(dev), ids_b.to(dev), 20); torch.cuda.synchronize()
report("rrf_fuse", bool((gi.cpu() == ri).all()) and bool(torch.allclose(gs.cpu()[ri >= 0], rs[ri >= 0], atol=1e-6)))
tok_store = torch.randint(1000, 5000, (60, 24), dtype=torch.int32); len_store = torch.randint(0, 25, (60,), dtype=torch.int32)
q_tok = torch.randint(1000, 5000, (64, 8), dtype=torch.int32); q_len = torch.randint(1, 9, (64,), dtype=torch.int32)
rp, rl = F.build_pairs_ref(q_tok, q_len, ri, tok_store, len_store, 24, 48)
ts, ls = tok_store.to(dev), len_store.to(dev)
gp, gl = F.build_pairs(q_tok.to(dev), q_len.to(dev), gi, F.ptr_table([ts], dev), F.ptr_table([ls], dev), 60, 24, 48)
torch.cuda.synchronize()
report("build_pairs", bool((gp.cpu() == rp).all()) and bool((gl.cpu() == rl).all()))
logits = torch.randn(64, 20); logits[2, 3] = float("nan")
ss, si = F.rerank_select(logits.to(dev), gi, 10); torch.cuda.synchronize()
ok = True
for q in range(64):
    l = logits[q].clone(); ids = ri[q]
    l[(ids < 0) | torch.isnan(l)] = float("-inf")
    order = sorted(range(20), key=lambda c: (-l[c].item(), c))[:10]
    ok &= [int(ids[c]) for c in order] == si[q].cpu().tolist()
report("rerank_select", ok)
print("MISC_ALL_OK" if ok_all else "MISC_FAIL")
