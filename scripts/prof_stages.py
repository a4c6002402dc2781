"""Per-stage CUDA-event timing of the hybrid pipeline on ONE GPU, emulating rank 0 of a W-way job (shard of docs/W
documents, reranker on nq/W queries).  Shows where a multi-GPU step spends its time without profiling a multi-rank run.

usage: python scripts/prof_stages.py [W=8] [docs=10000000] [nq=64]"""
import sys

import torch

from infomesh_b200.engine.hybrid import HybridConfig, HybridEngine
from infomesh_b200.engine.synth import SynthConfig, SynthShard, make_queries
from infomesh_b200.ops import fuse as F

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
docs = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda:0")
n_local = docs // W
shard = SynthShard(SynthConfig(n_docs=n_local, n_docs_global=n_local, doc_base=0), device=dev)
cfg = HybridConfig(nq=nq, use_graph=False)
eng = HybridEngine(shard, cfg)
q_terms, q_tok, q_len, _ = make_queries(SynthConfig(n_docs=n_local, n_docs_global=n_local), nq, max_terms=cfg.max_terms,
                                        max_q_tokens=cfg.max_q_tokens, device=dev)
enc_ids = torch.randint(1000, 20000, (nq, cfg.enc_seq), dtype=torch.int32, device=dev)
enc_len = torch.full((nq,), 12, dtype=torch.int32, device=dev)
eng.load_inputs(enc_ids, enc_len, q_tok, q_len, q_terms)


def t(fn, n=20):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        out = fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, out


ms_enc, q_emb = t(eng._encode)
ms_dense, (de_s, de_i) = t(lambda: eng._dense_local(q_emb))
ms_bm, (bm_s, bm_i) = t(eng._bm25_local)
ms_fuse, (fu_s, fu_i) = t(lambda: eng._fuse(bm_i, de_i))
eng.nq_local = nq // W
ms_rr, logits = t(lambda: eng._rerank(fu_i))
tot = ms_enc + ms_dense + ms_bm + ms_fuse + ms_rr
print(f"W={W} docs/rank={n_local} nq={nq} (graph-replayed stages, ms): encode {ms_enc:.3f}  dense {ms_dense:.3f}  bm25 {ms_bm:.3f}  "
      f"fuse {ms_fuse:.3f}  rerank({nq // W}q x {cfg.n_rerank}) {ms_rr:.3f}  sum {tot:.3f}")
