import numpy as np, torch
from infomesh_b200.ops import bm25 as BM
from infomesh_b200.engine.synth import SynthConfig, SynthShard, make_queries
dev = "cuda"
cfg = SynthConfig(n_docs=50_000, n_docs_global=50_000, vocab_terms=5_000, doc_len=48, passage_len=64)
sh = SynthShard(cfg, device=dev, build_chunk=20_000)
csr = dict(off=sh.bm25.off.cpu().numpy(), doc=sh.bm25.doc.cpu().numpy(), tf=sh.bm25.tf.cpu().numpy(),
           doc_len=np.full(cfg.n_docs, cfg.doc_len, np.int32), df=sh.df_local.cpu().numpy())
qt, qtok, qlen, dids = make_queries(cfg, 4, device=dev)
print("terms", qt[:2].tolist())
ps, pi = sh.bm25.search_partials(qt.to(dev)); torch.cuda.synchronize()
print("partials shape", ps.shape, "valid per warp q0:", (pi[:, 0, :] >= 0).sum(1).tolist())
print("q0 warp0 top:", ps[0, 0, :5].tolist(), pi[0, 0, :5].tolist())
s, i = sh.bm25.search(qt.to(dev), k=20); torch.cuda.synchronize()
for q in range(2):
    ref = BM.bm25_ref(csr, qt[q].tolist(), k=20)
    print("ref", ref[:5]); print("got", list(zip(s[q, :5].tolist(), i[q, :5].tolist())))
    print("idf", [float(sh.bm25.idf[t]) for t in qt[q].tolist() if t >= 0], "df", [int(csr['df'][t]) for t in qt[q].tolist() if t >= 0])
