"""Packed (varlen) cross-encoder forward for profiling: `ncu --metrics gpu__time_duration.sum ... python scripts/prof_reranker_packed.py mxfp8`.
Pair lengths mimic the benchmark batch (1280 pairs, mean ~70 tokens)."""
import sys

import torch

sys.path.insert(0, ".")
from infomesh_b200.models.bert import BGE_RERANKER_BASE, BertModel  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "mxfp8"
B, S = 1280, 128
m = BertModel(BGE_RERANKER_BASE, device="cuda", seed=1)
g = torch.Generator().manual_seed(0)
ids = torch.randint(5, 200000, (B, S), generator=g, dtype=torch.int32).cuda()
lens = torch.randint(66, 76, (B,), generator=g, dtype=torch.int32).cuda()
for _ in range(3):
    out = m.score_packed(ids, lens, precision=prec)
torch.cuda.synchronize()
if len(sys.argv) > 2:       # timing mode (no profiler): back-to-back forwards
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m.score_packed(ids, lens, precision=prec)
    e1.record()
    torch.cuda.synchronize()
    print(prec, "ms per forward:", e0.elapsed_time(e1) / 20)
