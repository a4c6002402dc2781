"""World-size-1 A/B of the fused TP/SP path against the plain model: isolates the kernel-side cost of the push
epilogue / flag waits from NVLink effects.  Also a convenient single-process target for `ncu`."""
import sys
from dataclasses import replace

import torch

from infomesh_b200.models.bert import BGE_RERANKER_BASE, BertModel
from infomesh_b200.parallel.tp import TPBertModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
S = 128
dev = torch.device("cuda:0")
cfg = replace(BGE_RERANKER_BASE, layers=layers)
ids = torch.randint(5, 5000, (B, S), dtype=torch.int32, device=dev)
lens = torch.full((B,), S, dtype=torch.int32, device=dev)
plain = BertModel(cfg, device=dev, seed=1)
tpm = TPBertModel(cfg, B, S, seed=1, comm="fused")


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


if len(sys.argv) > 3 and sys.argv[3] == "tp-only":
    for _ in range(3):
        tpm.score(ids, lens)
    torch.cuda.synchronize()
    sys.exit(0)
print(f"B={B} S={S} layers={layers}: plain {t(lambda: plain.score(ids, lens)):.3f} ms   fused-tp(world=1) {t(lambda: tpm.score(ids, lens)):.3f} ms")
